// integration/example_main.cc -- drives integration/batched_backend.h the way ceres::Solve would (EvaluationCallback::PrepareForEvaluation, then
// CostFunction::Evaluate per residual block) on a problem read from a flat file of doubles, and writes what Ceres would have received.
// tests/test_cpp_binding.py builds it with g++ against integration/ceres_stub and libkontiki_b200.so and compares the output with the Python binding.
//   usage: example_main IN.bin OUT.bin
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "batched_backend.h"

static std::vector<double> read_all(const char* path) {
  FILE* f = std::fopen(path, "rb");
  if (!f) { std::perror(path); std::exit(2); }
  std::fseek(f, 0, SEEK_END); const long bytes = std::ftell(f); std::fseek(f, 0, SEEK_SET);
  std::vector<double> v(bytes / sizeof(double));
  if (std::fread(v.data(), sizeof(double), v.size(), f) != v.size()) { std::fprintf(stderr, "short read\n"); std::exit(2); }
  std::fclose(f);
  return v;
}

int main(int argc, char** argv) {
  if (argc != 3) { std::fprintf(stderr, "usage: %s IN.bin OUT.bin\n", argv[0]); return 2; }
  const std::vector<double> in = read_all(argv[1]);
  size_t at = 0;
  auto take = [&](size_t n) { const double* p = in.data() + at; at += n; return p; };
  const double* h = take(18);
  const int n_knots = (int)h[0], ng = (int)h[3], nc = (int)h[4], n_lm = (int)h[5];
  const double dt = h[1], t0 = h[2];
  ktk_camera cam{};
  cam.base.q_ct[3] = 1.0; cam.base.max_time_offset = 0.1; cam.base.q_locked = cam.base.p_locked = cam.base.time_offset_locked = 1;
  cam.rows = (int)h[6]; cam.cols = (int)h[7]; cam.readout = h[8];
  for (int i = 0; i < 9; ++i) cam.K[i] = h[9 + i];
  cam.model = KTK_CAMERA_PINHOLE;
  // the reference keeps every knot in its own heap block and every inverse depth inside its Landmark
  std::vector<std::vector<double>> knot_store(n_knots, std::vector<double>(7));
  std::vector<double*> knot_blocks(n_knots);
  const double* knots = take((size_t)7 * n_knots);
  for (int k = 0; k < n_knots; ++k) { std::copy(knots + 7 * k, knots + 7 * k + 7, knot_store[k].begin()); knot_blocks[k] = knot_store[k].data(); }
  const double* gt = take(ng); const double* gy = take((size_t)3 * ng); const double* gw = take(ng);
  const double* ouv = take((size_t)2 * nc); const double* ot0 = take(nc); const double* ruv = take((size_t)2 * nc); const double* rt0 = take(nc);
  const double* lm = take(nc); const double* cw = take(nc); const double* hub = take(nc);
  const double* rho_in = take(n_lm);
  std::vector<double> rho(rho_in, rho_in + n_lm);

  try {
    kontiki::BatchedBackend backend(0);
    backend.SetSpline(dt, t0, knot_blocks);
    backend.SetCamera(cam);
    std::vector<int> grow(ng), crow(nc);
    for (int i = 0; i < ng; ++i) grow[i] = backend.AddGyroscopeRow(gt[i], gy + 3 * i, gw[i]);
    for (int i = 0; i < nc; ++i) crow[i] = backend.AddStaticRsRow(ouv + 2 * i, ot0[i], ruv + 2 * i, rt0[i], &rho[(size_t)lm[i]], cw[i], hub[i]);
    backend.Finalize();
    backend.PrepareForEvaluation(/*evaluate_jacobians=*/true, /*new_evaluation_point=*/true);

    std::vector<double> out;
    for (int i = 0; i < ng; ++i) {      // what ceres::Solve's residual-block loop does with a gyroscope block
      kontiki::BatchedImuCost cost(&backend, 0, grow[i]);
      double r[3], J[4][21], Jq[12], Jp[9], Jd[3];
      double* jac[7] = {J[0], J[1], J[2], J[3], Jq, Jp, Jd};
      cost.Evaluate(nullptr, r, jac);
      out.push_back((double)cost.first_knot());
      out.insert(out.end(), r, r + 3);
      for (int k = 0; k < 4; ++k) out.insert(out.end(), J[k], J[k] + 21);
    }
    const int cap = 24;
    std::vector<int32_t> ids((size_t)nc * cap), nids(nc);
    if (nc > 0 && ktk_get_structure(backend.problem(), backend.camera_group(), cap, ids.data(), nids.data()) != KTK_OK) throw std::runtime_error(ktk_last_error());
    for (int i = 0; i < nc; ++i) {      // ... and with a static-RS block: the structural knot list, the camera blocks, the landmark
      std::vector<int32_t> my(ids.begin() + (size_t)i * cap, ids.begin() + (size_t)i * cap + nids[i]);
      kontiki::BatchedStaticRsCost cost(&backend, crow[i], my);
      std::vector<std::vector<double>> Jk(my.size(), std::vector<double>(14));
      double r[2], Jq[8], Jp[6], Jd[2], Jrho[2];
      std::vector<double*> jac;
      for (auto& b : Jk) jac.push_back(b.data());
      jac.push_back(Jq); jac.push_back(Jp); jac.push_back(Jd); jac.push_back(Jrho);
      cost.Evaluate(nullptr, r, jac.data());
      out.push_back((double)nids[i]);
      for (int c = 0; c < cap; ++c) out.push_back(c < nids[i] ? (double)my[c] : -1.0);
      out.insert(out.end(), r, r + 2);
      for (int c = 0; c < cap; ++c) for (int e = 0; e < 14; ++e) out.push_back(c < nids[i] ? Jk[c][e] : 0.0);
      out.insert(out.end(), Jrho, Jrho + 2);
    }
    FILE* f = std::fopen(argv[2], "wb");
    if (!f) { std::perror(argv[2]); return 2; }
    std::fwrite(out.data(), sizeof(double), out.size(), f);
    std::fclose(f);
  } catch (const std::exception& e) {
    std::fprintf(stderr, "error: %s\n", e.what());
    return 1;
  }
  return 0;
}
