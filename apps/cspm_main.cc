// apps/cspm_main.cc — the reference's main() flow (CSPM/main.cc:57-139) on top
// of include/cspm.hpp, with the same flag names (gflags syntax --name=value)
// and the same constants (max_iter=3, wnd_size=35, scale_num=5).  Image I/O is
// binary PPM (P6) in / PGM (P5) out: OpenCV's imread/imwrite are not available
// here and image decoding is outside the hot path.  --synth=WxHxD[:seed] generates the pair in memory with the
// dependency-free generator of include/cspm_synth.h instead (the benchmark's inputs: --synth=1920x1080x256:0xC5B20003
// is BASELINE.json's HD pair), in which case max_dis defaults to D and dis_scale to 1.  Extra flags beside the
// reference's: --scale_num (main.cc:100 hard-codes 5), --schedule=raster|redblack, --seed, --device.
//
// build: g++ -O2 -std=c++17 -Iinclude apps/cspm_main.cc -o apps/cspm_main \
//            -Lcrossscalepatchmatch_b200 -lcspm_b200 -Wl,-rpath,'$ORIGIN/../crossscalepatchmatch_b200'
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <map>
#include <string>

#include "cspm.hpp"
#include "cspm_synth.h"

using namespace cspm;

static bool read_ppm(const std::string& path, Mat* bgr) {
  std::ifstream f(path, std::ios::binary);
  std::string magic;
  int w, h, mx;
  if (!(f >> magic >> w >> h >> mx) || magic != "P6" || mx != 255) return false;
  f.get();
  *bgr = Mat(h, w, CV_8UC3);
  std::vector<uint8_t> row(3 * (size_t)w);
  for (int y = 0; y < h; ++y) {
    if (!f.read((char*)row.data(), row.size())) return false;
    uint8_t* d = bgr->ptr<uint8_t>(y);
    for (int x = 0; x < w; ++x) {  // RGB file order -> BGR memory order (what imread returns)
      d[3 * x] = row[3 * x + 2]; d[3 * x + 1] = row[3 * x + 1]; d[3 * x + 2] = row[3 * x];
    }
  }
  return true;
}
static bool write_pgm(const std::string& path, const Mat& m) {
  std::ofstream f(path, std::ios::binary);
  f << "P5\n" << m.cols << " " << m.rows << "\n255\n";
  for (int y = 0; y < m.rows; ++y) f.write((const char*)m.ptr<uint8_t>(y), m.cols);
  return (bool)f;
}

int main(int argc, char** argv) {
  std::cout << "PatchMatch Stereo Matching" << std::endl;
  // flag defaults of main.cc:23-34
  std::map<std::string, std::string> F = {{"l_img_file", "l_img.ppm"}, {"r_img_file", "r_img.ppm"},
                                          {"l_dis_file", "l_dis.pgm"}, {"r_dis_file", "r_dis.pgm"},
                                          {"max_dis", "0"}, {"dis_scale", "0"}, {"cc_name", "CCName"},
                                          {"use_cs", "false"}, {"use_pp", "false"}, {"reg_lambda", "0.0"},
                                          {"seed", "24301"}, {"device", "0"}, {"synth", ""}, {"scale_num", "5"},
                                          {"schedule", "raster"}};
  for (int i = 1; i < argc; ++i) {
    std::string a = argv[i];
    while (!a.empty() && a[0] == '-') a.erase(0, 1);
    size_t eq = a.find('=');
    std::string k = a.substr(0, eq), v = eq == std::string::npos ? "true" : a.substr(eq + 1);
    if (k.rfind("no", 0) == 0 && F.count(k.substr(2))) { k = k.substr(2); v = "false"; }
    if (!F.count(k)) { std::cerr << "unknown flag --" << k << std::endl; return EXIT_FAILURE; }
    F[k] = v;
  }
  auto flag_true = [&](const char* k) { return F[k] == "true" || F[k] == "1"; };
  Mat l_img, r_img;
  if (!F["synth"].empty()) {
    int w = 0, h = 0, d = 0;
    unsigned long long seed = CSPM_SYNTH_SEED_BASE;
    char tail[64] = "";
    if (std::sscanf(F["synth"].c_str(), "%dx%dx%d%63s", &w, &h, &d, tail) < 3 || w <= 0 || h <= 0 || d <= 0) {
      std::cerr << "--synth expects WxHxD[:seed]" << std::endl;
      return EXIT_FAILURE;
    }
    if (tail[0] == ':') seed = std::strtoull(tail + 1, nullptr, 0);
    l_img = Mat(h, w, CV_8UC3);
    r_img = Mat(h, w, CV_8UC3);
    if (cspm_synth_pair(w, h, d, seed, l_img.data, r_img.data, nullptr) != 0) return EXIT_FAILURE;
    if (F["max_dis"] == "0") F["max_dis"] = std::to_string(d);
    if (F["dis_scale"] == "0") F["dis_scale"] = "1";
  } else if (!read_ppm(F["l_img_file"], &l_img) || !read_ppm(F["r_img_file"], &r_img)) {
    std::cout << "Error: can not open image\n";  // main.cc:70-75
    return EXIT_FAILURE;
  }
  try {
    const int device = std::atoi(F["device"].c_str());
    CCMethod* cc_cost = GetCCType(F["cc_name"]);  // main.cc:89
    auto t0 = std::chrono::steady_clock::now();   // timed region starts, main.cc:92
    const int max_iter = 3, wnd_size = 35;
    IPlaneCost* plane_cost = nullptr;
    if (flag_true("use_cs")) {
      const int scale_num = std::atoi(F["scale_num"].c_str());  // main.cc:100 (5 there)
      plane_cost = new PreCSPC(l_img, r_img, std::atoi(F["max_dis"].c_str()), wnd_size, scale_num, cc_cost,
                               std::atof(F["reg_lambda"].c_str()), device);
    } else {
      plane_cost = new PreSSPC(l_img, r_img, std::atoi(F["max_dis"].c_str()), wnd_size, cc_cost, device);
    }
    CSPatchMatch* patch_match = new CSPatchMatch(l_img, r_img, std::atoi(F["max_dis"].c_str()),
                                                 std::atoi(F["dis_scale"].c_str()),
                                                 std::strtoull(F["seed"].c_str(), nullptr, 0),
                                                 F["schedule"] == "redblack" ? CSPM_SCHEDULE_REDBLACK : CSPM_SCHEDULE_RASTER, device);
    patch_match->PatchMatch(max_iter, plane_cost, flag_true("use_pp"));
    double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    std::cout << "--------------------------------------------------------\n";
    std::cout << "Total Time: " << sec << std::endl;  // main.cc:122-125
    std::cout << "Mpixel*disparities/s: " << (double)l_img.cols * l_img.rows * std::atoi(F["max_dis"].c_str()) / sec / 1e6 << std::endl;
    std::cout << "--------------------------------------------------------\n";
    write_pgm(F["l_dis_file"], patch_match->dis(kLeft));
    write_pgm(F["r_dis_file"], patch_match->dis(kRight));
    delete patch_match;
    delete plane_cost;
    delete cc_cost;
  } catch (const Exception& e) {
    std::cerr << "cspm error " << e.code << ": " << e.what() << std::endl;
    return EXIT_FAILURE;
  }
  return EXIT_SUCCESS;
}
