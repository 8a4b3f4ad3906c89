// Drop-in replacement for the reference's libsuper4pcs.so entry point: exports the SAME mangled
// C++ symbol the ROS node resolves (declared at PPE/src/hypothesis_generation/
// ObjectPoseCandidateSet.cpp:5-9, defined at S4/super4pcs_test.cc:39-111, linked through
// PPE/CMakeLists.txt:116), reads the same files, fills the same outputs -- and does all the work on
// the B200 through the C ABI of include/pgp.h.  Host-only C++; no CUDA, no Eigen, no PCL needed.
//
//   in : segment / model_validation / model_search PLY paths (as PCL's savePLYFile writes them),
//        16-bit probability PNG path, PPFMap (uploaded to the device; selects the shipped operMode 1 = StoCS base
//        sampling + PPF-map pair lookup; an EMPTY map selects operMode 0 = wide random bases + device-side pair
//        extraction), intrinsics, object name, scene path
//   out: bestHypothesis (pose, score), hypothesisSet = the strictly-improving chain in generation
//        order (match4pcsBase.cc:1888-1914), registered_points (scene indices matched by the best pose)
// Failure behaviour: never throws across the boundary; on any error the outputs are identity / 0 /
// empty (the reference: exit(-1) on unreadable input, uninitialised outputs on exceptions).
//
// A long-lived service component, like the node it plugs into (one process, many requests):
//   * one device group for the life of the process (pgp_group_*): PGP_DEVICES = "0,1,2,3" shards every request's bases over
//     those GPUs -- each generates and scores its own hypotheses, the selections are all-gathered over NCCL and merged inside
//     libpgp.so -- with results identical to the single-GPU ones; default: the one device PGP_DEVICE names (0);
//   * per-object caches: GlobalCfg::loadObjects loads every model ONCE (PPE/src/data_layer/GlobalCfg.cpp:30-64) and then re-writes
//     the same two PLYs on every request (ObjectPoseCandidateSet.cpp:53-60); here a model slot is keyed by the object name and
//     validated by a hash of the raw bytes of both files (no ASCII parse on a hit) and a digest of the caller's PPFMap, so parsing,
//     pgp_set_model and the PPF-map upload happen once per object, not once per request;
//   * a mutex around the shared group (the commented-out per-object threads of SceneCfg.cpp:377,404-405 may be revived);
//   * the reference's one file side effect: <scenePath>debug_super4PCS/<objName>_time.txt gets one line per pose of the returned
//     chain -- the time at which that pose became the best (match4pcsBase.cc:1896,1909-1913); here every pose of the chain is
//     available at the same instant, the end of the request, so that is the value written.
//
// Environment: PGP_DEVICES / PGP_DEVICE (see above), PGP_LCP_MODE = weighted (default, the shipped
// WeightedVerify) | count, PGP_DELTA (default 0.005 = S4/super4pcs_test.cc:20), PGP_SEED, PGP_PCS_MODE = stocs | super4pcs | v4pcs
// (operMode 1 / 0 / 2; default: stocs when the caller's PPFMap is not empty -- the reference hard-sets operMode 1,
// match4pcsBase.cc:300; operMode 2 is scored with Verify like the reference does, :1498-1499).
#include <zlib.h>

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <chrono>
#include <map>
#include <mutex>
#include <sstream>
#include <string>
#include <utility>
#include <vector>

#include "eigen_abi.h"
#include "pgp.h"

namespace {

struct Cloud { std::vector<float> xyz, nrm; };

// ---- PLY: vertex element with named properties, ascii or binary_little_endian; other elements skipped
bool read_ply(const std::string& path, Cloud& out) {
  std::ifstream in(path, std::ios::binary);
  if (!in) return false;
  std::string tok;
  in >> tok;
  if (tok != "ply") return false;
  struct Prop { std::string type, name; };
  std::vector<Prop> props;
  bool ascii = true, in_vertex = false;
  size_t n_vertex = 0;
  std::string line;
  std::getline(in, line);
  while (std::getline(in, line)) {
    if (!line.empty() && line.back() == '\r') line.pop_back();
    std::istringstream ls(line);
    std::string key;
    ls >> key;
    if (key == "end_header") break;
    if (key == "format") { std::string f; ls >> f; if (f == "ascii") ascii = true; else if (f == "binary_little_endian") ascii = false; else return false; }
    else if (key == "element") { std::string name; size_t cnt; ls >> name >> cnt; in_vertex = (name == "vertex"); if (in_vertex) n_vertex = cnt; }
    else if (key == "property" && in_vertex) { Prop p; ls >> p.type; if (p.type == "list") return false; ls >> p.name; props.push_back(p); }
  }
  auto find = [&](std::initializer_list<const char*> names) { for (size_t i = 0; i < props.size(); ++i) for (const char* n : names) if (props[i].name == n) return (int)i; return -1; };
  const int ix = find({"x"}), iy = find({"y"}), iz = find({"z"});
  const int inx = find({"nx", "normal_x"}), iny = find({"ny", "normal_y"}), inz = find({"nz", "normal_z"});
  if (ix < 0 || iy < 0 || iz < 0) return false;
  const bool has_n = inx >= 0 && iny >= 0 && inz >= 0;
  out.xyz.resize(n_vertex * 3);
  out.nrm.assign(has_n ? n_vertex * 3 : 0, 0.f);
  auto size_of = [](const std::string& t) -> int {
    if (t == "float" || t == "float32" || t == "int" || t == "int32" || t == "uint" || t == "uint32") return 4;
    if (t == "double" || t == "float64") return 8;
    if (t == "uchar" || t == "uint8" || t == "char" || t == "int8") return 1;
    if (t == "short" || t == "int16" || t == "ushort" || t == "uint16") return 2;
    return -1;
  };
  std::vector<double> v(props.size());
  for (size_t i = 0; i < n_vertex; ++i) {
    if (ascii) {
      for (size_t k = 0; k < props.size(); ++k) if (!(in >> v[k])) return false;
    } else {
      for (size_t k = 0; k < props.size(); ++k) {
        const int sz = size_of(props[k].type);
        char b[8];
        if (sz < 0 || !in.read(b, sz)) return false;
        const std::string& t = props[k].type;
        if (t == "float" || t == "float32") { float f; memcpy(&f, b, 4); v[k] = f; }
        else if (t == "double" || t == "float64") { double d; memcpy(&d, b, 8); v[k] = d; }
        else if (sz == 1) v[k] = (unsigned char)b[0];
        else if (sz == 2) { uint16_t u; memcpy(&u, b, 2); v[k] = u; }
        else { int32_t q; memcpy(&q, b, 4); v[k] = q; }
      }
    }
    out.xyz[3 * i] = (float)v[ix]; out.xyz[3 * i + 1] = (float)v[iy]; out.xyz[3 * i + 2] = (float)v[iz];
    if (has_n) { out.nrm[3 * i] = (float)v[inx]; out.nrm[3 * i + 1] = (float)v[iny]; out.nrm[3 * i + 2] = (float)v[inz]; }
  }
  return n_vertex > 0;
}

// ---- PNG: 8/16-bit grayscale, non-interlaced (what cv::imwrite produces for the CV_16UC1 prior image)
bool read_png_gray16(const std::string& path, std::vector<uint16_t>& img, int& rows, int& cols) {
  std::ifstream in(path, std::ios::binary);
  if (!in) return false;
  std::vector<unsigned char> f((std::istreambuf_iterator<char>(in)), std::istreambuf_iterator<char>());
  static const unsigned char sig[8] = {137, 80, 78, 71, 13, 10, 26, 10};
  if (f.size() < 8 || memcmp(f.data(), sig, 8)) return false;
  auto be32 = [&](size_t o) { return (uint32_t)f[o] << 24 | (uint32_t)f[o + 1] << 16 | (uint32_t)f[o + 2] << 8 | f[o + 3]; };
  size_t o = 8;
  int depth = 0, ctype = -1, interlace = 0;
  std::vector<unsigned char> z;
  while (o + 8 <= f.size()) {
    const uint32_t len = be32(o);
    const std::string type((const char*)&f[o + 4], 4);
    if (o + 12 + len > f.size()) return false;
    if (type == "IHDR") { cols = (int)be32(o + 8); rows = (int)be32(o + 12); depth = f[o + 16]; ctype = f[o + 17]; interlace = f[o + 20]; }
    else if (type == "IDAT") z.insert(z.end(), f.begin() + o + 8, f.begin() + o + 8 + len);
    else if (type == "IEND") break;
    o += 12 + len;
  }
  if (ctype != 0 || interlace != 0 || (depth != 8 && depth != 16) || rows <= 0 || cols <= 0) return false;
  const size_t bpp = depth / 8, stride = (size_t)cols * bpp;
  std::vector<unsigned char> raw((stride + 1) * rows);
  uLongf rl = raw.size();
  if (uncompress(raw.data(), &rl, z.data(), z.size()) != Z_OK || rl != raw.size()) return false;
  std::vector<unsigned char> cur(stride), prev(stride, 0);
  img.resize((size_t)rows * cols);
  for (int r = 0; r < rows; ++r) {
    const unsigned char* line = &raw[(stride + 1) * r];
    const int ft = line[0];
    for (size_t i = 0; i < stride; ++i) {
      const int a = i >= bpp ? cur[i - bpp] : 0, b = prev[i], c = i >= bpp ? prev[i - bpp] : 0;
      int x = line[1 + i];
      switch (ft) {
        case 1: x += a; break;
        case 2: x += b; break;
        case 3: x += (a + b) / 2; break;
        case 4: { const int p = a + b - c, pa = std::abs(p - a), pb = std::abs(p - b), pc = std::abs(p - c); x += (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c); break; }
        default: break;
      }
      cur[i] = (unsigned char)x;
    }
    for (int cidx = 0; cidx < cols; ++cidx)
      img[(size_t)r * cols + cidx] = depth == 16 ? (uint16_t)(cur[2 * cidx] << 8 | cur[2 * cidx + 1]) : (uint16_t)cur[cidx];
    prev = cur;
  }
  return true;
}

// FNV-1a over the raw bytes of a file (0 when unreadable): the cache key of a model's PLY
uint64_t file_hash(const std::string& path) {
  std::ifstream in(path, std::ios::binary);
  if (!in) return 0;
  uint64_t h = 1469598103934665603ull;
  char buf[1 << 16];
  while (in.read(buf, sizeof(buf)) || in.gcount() > 0) {
    const std::streamsize n = in.gcount();
    for (std::streamsize i = 0; i < n; ++i) { h ^= (unsigned char)buf[i]; h *= 1099511628211ull; }
    if (n < (std::streamsize)sizeof(buf)) break;
  }
  return h ? h : 1;
}

// digest of the caller's PPFMap (it is loaded once per object from PPFMap.txt and handed over by reference on every request)
uint64_t ppf_digest(const std::map<std::vector<int>, std::vector<std::pair<int, int>>>& m) {
  uint64_t h = 1469598103934665603ull ^ (uint64_t)m.size();
  auto mix = [&](uint64_t v) { h ^= v; h *= 1099511628211ull; };
  size_t k = 0;
  const size_t step = m.size() / 64 + 1;           // every row's size, 64 rows' content
  for (const auto& kv : m) {
    mix(kv.second.size());
    if (k++ % step == 0) {
      for (int c : kv.first) mix((uint64_t)(uint32_t)c);
      if (!kv.second.empty()) { mix((uint64_t)(uint32_t)kv.second.front().first << 32 | (uint32_t)kv.second.front().second); mix((uint64_t)(uint32_t)kv.second.back().first << 32 | (uint32_t)kv.second.back().second); }
    }
  }
  return h;
}

struct ModelSlot {
  std::string name;
  uint64_t val_hash = 0, search_hash = 0, ppf = 0;   // what is resident on the devices for this slot
  int nv = 0, nq = 0;
  bool has_nrm = false, ppf_built = false;
  uint64_t last_use = 0;
};

struct Service {
  pgp_group* g = nullptr;
  std::mutex mu;
  std::vector<ModelSlot> slots;
  uint64_t tick = 0;
  uint64_t n_requests = 0, n_model_hits = 0;
  ~Service() { if (g) pgp_group_destroy(g); }
};
Service& service() {
  static Service s;
  return s;
}
// creates the device group on first use: PGP_DEVICES = comma-separated CUDA devices, else PGP_DEVICE, else device 0
pgp_group* shared_group(Service& sv) {
  if (!sv.g) {
    std::vector<int> ids;
    if (const char* ds = getenv("PGP_DEVICES")) {
      std::stringstream ss(ds);
      std::string tok;
      while (std::getline(ss, tok, ',')) if (!tok.empty()) ids.push_back(atoi(tok.c_str()));
    }
    if (ids.empty()) { const char* d = getenv("PGP_DEVICE"); ids.push_back(d ? atoi(d) : 0); }
    sv.g = pgp_group_create((int)ids.size(), ids.data());
    if (!sv.g) fprintf(stderr, "[pgp] %s\n", pgp_last_error(nullptr));
    sv.slots.resize(PGP_MAX_OBJECTS);
  }
  return sv.g;
}
// the model slot of `name`: an existing one, else a free one, else the least recently used
int slot_of(Service& sv, const std::string& name) {
  int lru = 0;
  for (int i = 0; i < (int)sv.slots.size(); ++i) if (sv.slots[i].name == name && sv.slots[i].val_hash) return i;
  for (int i = 0; i < (int)sv.slots.size(); ++i) {
    if (!sv.slots[i].val_hash) return i;
    if (sv.slots[i].last_use < sv.slots[lru].last_use) lru = i;
  }
  sv.slots[lru] = ModelSlot();
  return lru;
}

void to_isometry(const double* P16, Eigen::Isometry3d& iso) {
#ifdef PGP_USE_REAL_EIGEN
  for (int r = 0; r < 4; ++r) for (int c = 0; c < 4; ++c) iso.matrix()(r, c) = P16[4 * r + c];
#else
  for (int r = 0; r < 4; ++r) for (int c = 0; c < 4; ++c) iso(r, c) = P16[4 * r + c];
#endif
}

}  // namespace

// NB: C++ linkage on purpose -- this must mangle exactly like the reference's definition.
void getProbableTransformsSuper4PCS(std::string input1, std::string input2, std::string input3,
                                    std::pair<Eigen::Isometry3d, float>& bestHypothesis,
                                    std::vector<std::pair<Eigen::Isometry3d, float>>& hypothesisSet,
                                    std::string probImagePath,
                                    std::map<std::vector<int>, std::vector<std::pair<int, int>>>& PPFMap,
                                    int max_count_ppf, Eigen::Matrix3f camIntrinsic, std::string objName, std::string scenePath,
                                    std::vector<int>& registered_points) __attribute__((visibility("default")));

void getProbableTransformsSuper4PCS(std::string input1, std::string input2, std::string input3,
                                    std::pair<Eigen::Isometry3d, float>& bestHypothesis,
                                    std::vector<std::pair<Eigen::Isometry3d, float>>& hypothesisSet,
                                    std::string probImagePath,
                                    std::map<std::vector<int>, std::vector<std::pair<int, int>>>& PPFMap,
                                    int max_count_ppf, Eigen::Matrix3f camIntrinsic, std::string objName, std::string scenePath,
                                    std::vector<int>& registered_points) {
  (void)max_count_ppf;
  const auto t_start = std::chrono::steady_clock::now();
  Eigen::Isometry3d identity;
#ifdef PGP_USE_REAL_EIGEN
  identity.setIdentity();
#endif
  bestHypothesis.first = identity;                  // "no pose" -> identity + 0 (match4pcsBase.cc:1791-1796)
  bestHypothesis.second = 0.f;
  hypothesisSet.clear();                            // Perform_N_steps: allPose.clear() before the chain is pushed (:1903)
  registered_points.clear();
  try {
    Service& sv = service();
    std::lock_guard<std::mutex> lock(sv.mu);
    pgp_group* g = shared_group(sv);
    if (!g) return;
    pgp_ctx* ctx0 = pgp_group_ctx(g, 0);
    sv.n_requests++;
    auto fail = [&](const char* what) { fprintf(stderr, "[pgp] %s: %s\n", what, pgp_group_last_error(g)); };
    // argument order of the reference: set1 = segment (P), set2 = input2 = model_validation copy (Q_validation),
    // set3 = input3 = model_search copy (Q)   (S4/super4pcs_test.cc:58-74,103)
    Cloud seg;
    if (!read_ply(input1, seg)) { fprintf(stderr, "[pgp] cannot read the segment PLY %s\n", input1.c_str()); return; }
    const char* de = getenv("PGP_DELTA");
    const float delta = de ? (float)atof(de) : 0.005f;
    const char* me = getenv("PGP_LCP_MODE");
    int mode = (me && !strcmp(me, "count")) ? PGP_LCP_COUNT : PGP_LCP_WEIGHTED;
    const char* se = getenv("PGP_SEED");
    const uint64_t seed = se ? strtoull(se, nullptr, 10) : 1;
    const int ns = (int)(seg.xyz.size() / 3);

    // ---- the object's model slot: parse + upload only when the files' bytes changed
    const uint64_t hv = file_hash(input2), hs = file_hash(input3);
    if (!hv || !hs) { fprintf(stderr, "[pgp] cannot read the model PLY files\n"); return; }
    const int obj = slot_of(sv, objName);
    ModelSlot& ms = sv.slots[obj];
    ms.last_use = ++sv.tick;
    if (ms.name != objName || ms.val_hash != hv || ms.search_hash != hs) {
      Cloud val, search;
      if (!read_ply(input2, val) || !read_ply(input3, search)) { fprintf(stderr, "[pgp] cannot parse the model PLY files\n"); return; }
      const int nv = (int)(val.xyz.size() / 3), nq = (int)(search.xyz.size() / 3);
      ms = ModelSlot();
      if (pgp_group_set_model(g, obj, search.xyz.data(), search.nrm.empty() ? nullptr : search.nrm.data(), nq, val.xyz.data(),
                              val.nrm.empty() ? nullptr : val.nrm.data(), nv))
        return fail("set_model");
      ms.name = objName; ms.val_hash = hv; ms.search_hash = hs; ms.nv = nv; ms.nq = nq;
      ms.has_nrm = !val.nrm.empty() && !search.nrm.empty();
      ms.last_use = sv.tick;
    } else {
      sv.n_model_hits++;
    }
    if (seg.nrm.empty() || !ms.has_nrm) mode = PGP_LCP_COUNT;          // the normal gate needs normals on both sides

    if (pgp_group_set_scene(g, seg.xyz.data(), seg.nrm.empty() ? nullptr : seg.nrm.data(), ns, delta)) return fail("set_scene");
    std::vector<uint16_t> img;
    int rows = 0, cols = 0;
    if (mode == PGP_LCP_WEIGHTED && !probImagePath.empty() && read_png_gray16(probImagePath, img, rows, cols)) {
      float K[9];
      for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) K[3 * r + c] = camIntrinsic(r, c);
      if (pgp_group_set_scene_prior_image(g, img.data(), rows, cols, K)) return fail("set_scene_prior_image");
    }
    pgp_pcs_opts opts;
    pgp_pcs_default_opts(&opts);                    // 100 bases x <= 100 congruent quads (match4pcsBase.cc:290,1858)
    const char* pe = getenv("PGP_PCS_MODE");
    bool stocs = !PPFMap.empty() && !seg.nrm.empty() && ms.has_nrm;
    const bool v4pcs = pe && !strcmp(pe, "v4pcs");
    if (pe && (!strcmp(pe, "super4pcs") || v4pcs)) stocs = false;
    if (pe && !strcmp(pe, "stocs") && PPFMap.empty()) {
      if (!ms.ppf_built) {
        if (pgp_group_build_ppf_map(g, obj)) return fail("build_ppf_map");   // no PPFMap.txt was loaded: build the map from the search cloud, once
        ms.ppf_built = true; ms.ppf = 0;
      }
      stocs = !seg.nrm.empty() && ms.has_nrm;
    } else if (stocs) {
      const uint64_t pd = ppf_digest(PPFMap);
      if (ms.ppf != pd || ms.ppf_built) {
        // std::map<vector<int>, vector<pair<int,int>>> (Objects::readPPFMap, PPE/src/data_layer/Objects.cpp:31-49) -> flat rows
        std::vector<int32_t> keys4, prs;
        std::vector<int64_t> offs;
        keys4.reserve(PPFMap.size() * 4);
        offs.reserve(PPFMap.size() + 1);
        for (const auto& kv : PPFMap) {
          if (kv.first.size() != 4) continue;
          offs.push_back((int64_t)prs.size() / 2);
          for (int c = 0; c < 4; ++c) keys4.push_back(kv.first[c]);
          for (const auto& pr : kv.second) { prs.push_back(pr.first); prs.push_back(pr.second); }
        }
        offs.push_back((int64_t)prs.size() / 2);
        if (pgp_group_set_ppf_map(g, obj, keys4.data(), offs.data(), prs.data(), (int64_t)keys4.size() / 4)) return fail("set_ppf_map");
        ms.ppf = pd; ms.ppf_built = false;
      }
    }
    opts.mode = v4pcs ? 2 : stocs ? 1 : 0;
    if (v4pcs) mode = PGP_LCP_COUNT;                // verifyRigidTransform uses Verify in operMode 2 (match4pcsBase.cc:1498-1499)
    int64_t n_hyp = 0;
    if (pgp_group_generate_pcs(g, obj, &opts, seed, 10000, &n_hyp)) return fail("generate_pcs");
    if (n_hyp == 0) return;
    if (pgp_group_score_generated(g, obj, mode)) return fail("score_generated");
    std::vector<pgp_hyp> chain(4096);
    int n_chain = pgp_group_improving_chain(g, obj, chain.data(), (int)chain.size());
    if (n_chain < 0) { fail("improving_chain"); return; }
    if (n_chain == 0) return;
    for (int i = 0; i < n_chain; ++i) {
      double P[16];
      if (pgp_centred_to_pose(ctx0, obj, chain[i].T, P)) { fprintf(stderr, "[pgp] centred_to_pose: %s\n", pgp_last_error(ctx0)); hypothesisSet.clear(); return; }
      std::pair<Eigen::Isometry3d, float> e;
      to_isometry(P, e.first);
      e.second = chain[i].score;
      hypothesisSet.push_back(e);
    }
    bestHypothesis = hypothesisSet.back();          // == allPose[best_lcp_index], best_LCP_
    if (mode == PGP_LCP_WEIGHTED) {
      registered_points.resize(ms.nv);
      const int k = pgp_registered_points(ctx0, obj, chain[n_chain - 1].T, registered_points.data(), ms.nv);
      registered_points.resize(k > 0 ? k : 0);
    }
    // the reference's per-object time log (match4pcsBase.cc:1909-1913; opened in append mode, silently skipped when the
    // directory does not exist -- the caller creates it, SceneCfg.cpp:323-331)
    {
      const float t = std::chrono::duration<float>(std::chrono::steady_clock::now() - t_start).count();
      std::ofstream tf((scenePath + "debug_super4PCS/" + objName + "_time.txt").c_str(), std::ofstream::out | std::ofstream::app);
      if (tf) for (int i = 0; i < n_chain; ++i) tf << t << std::endl;
    }
    if (getenv("PGP_VERBOSE"))
      fprintf(stderr, "[pgp] %s: %lld hypotheses on %d device(s), chain %d, best %.4f, %.2f ms (requests %llu, model-cache hits %llu)\n", objName.c_str(),
              (long long)n_hyp, pgp_group_size(g), n_chain, bestHypothesis.second,
              std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t_start).count(), (unsigned long long)sv.n_requests,
              (unsigned long long)sv.n_model_hits);
  } catch (...) {
    // swallow everything, like S4/super4pcs_test.cc:101-108 -- but with defined outputs
    hypothesisSet.clear();
    registered_points.clear();
    bestHypothesis.first = identity;
    bestHypothesis.second = 0.f;
  }
}
