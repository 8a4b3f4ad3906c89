// Test driver: calls the drop-in exactly as CongruentSetMatching::generate does
// (PPE/src/hypothesis_generation/ObjectPoseCandidateSet.cpp:66-68) and prints the outputs as JSON.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <map>
#include <string>
#include <utility>
#include <vector>

#include "eigen_abi.h"

void getProbableTransformsSuper4PCS(std::string input1, std::string input2, std::string input3,
                                    std::pair<Eigen::Isometry3d, float>& bestHypothesis,
                                    std::vector<std::pair<Eigen::Isometry3d, float>>& hypothesisSet,
                                    std::string probImagePath,
                                    std::map<std::vector<int>, std::vector<std::pair<int, int>>>& PPFMap,
                                    int max_count_ppf, Eigen::Matrix3f camIntrinsic, std::string objName, std::string scenePath,
                                    std::vector<int>& registered_points);

int main(int argc, char** argv) {
  if (argc < 4) { fprintf(stderr, "usage: %s segment.ply model_validation.ply model_search.ply [prob.png fx fy cx cy [PPFMap.txt]]\n", argv[0]); return 2; }
  std::pair<Eigen::Isometry3d, float> best;
  std::vector<std::pair<Eigen::Isometry3d, float>> set;
  std::map<std::vector<int>, std::vector<std::pair<int, int>>> ppf;
  std::vector<int> reg;
  Eigen::Matrix3f K;
  std::string png = argc > 4 ? argv[4] : "";
  if (argc > 8) { K(0, 0) = atof(argv[5]); K(1, 1) = atof(argv[6]); K(0, 2) = atof(argv[7]); K(1, 2) = atof(argv[8]); K(2, 2) = 1.f; }
  if (argc > 9) {   // PPFMap.txt, the format Objects::readPPFMap parses (PPE/src/data_layer/Objects.cpp:31-49): f1 f2 f3 f4 count, then count pairs
    std::ifstream f(argv[9]);
    std::vector<int> key(4);
    int cnt;
    while (f >> key[0] >> key[1] >> key[2] >> key[3] >> cnt) {
      std::vector<std::pair<int, int>> v;
      for (int i = 0; i < cnt; ++i) { int a, b; f >> a >> b; v.push_back(std::make_pair(a, b)); }
      ppf.insert(std::make_pair(key, v));
    }
  }
  // PGP_DRIVER_REPEAT = n: the same request n times in one process, as a long-lived node issues them (the caller's vector is NOT
  // emptied in between: the callee must clear it); PGP_DRIVER_SCENE: the scenePath argument (default /tmp/)
  const int repeat = getenv("PGP_DRIVER_REPEAT") ? atoi(getenv("PGP_DRIVER_REPEAT")) : 1;
  const std::string scene = getenv("PGP_DRIVER_SCENE") ? getenv("PGP_DRIVER_SCENE") : "/tmp/";
  std::vector<double> call_ms;
  for (int r = 0; r < repeat; ++r) {
    const auto t0 = std::chrono::steady_clock::now();
    getProbableTransformsSuper4PCS(argv[1], argv[2], argv[3], best, set, png, ppf, 0, K, "obj", scene, reg);
    call_ms.push_back(std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count());
  }
  printf("{\"call_ms\": [");
  for (size_t i = 0; i < call_ms.size(); ++i) printf("%s%.3f", i ? ", " : "", call_ms[i]);
  printf("], \"best_score\": %.9g, \"n_hypotheses\": %zu, \"n_registered\": %zu, \"best_pose\": [", best.second, set.size(), reg.size());
  for (int r = 0; r < 4; ++r) for (int c = 0; c < 4; ++c) printf("%s%.17g", (r || c) ? ", " : "", best.first(r, c));
  printf("], \"scores\": [");
  for (size_t i = 0; i < set.size(); ++i) printf("%s%.9g", i ? ", " : "", set[i].second);
  printf("], \"registered_head\": [");
  for (size_t i = 0; i < reg.size() && i < 8; ++i) printf("%s%d", i ? ", " : "", reg[i]);
  printf("]}\n");
  return 0;
}
