// TEST INFRASTRUCTURE ONLY -- never linked into, imported by, or called from the product path.
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs load it.
//
// liboracle.so: a plain CPU restatement of the AlayaLite hot path (batched k-NN search over the
// raw / SQ8 / SQ4 spaces, HNSW build, brute-force ground truth), written from the reference's
// behaviour, each function citing the reference file:line it follows (paths relative to
// /root/reference). Scalar, single-threaded per call, no SIMD: it is the checker, not a product.
//
// Parity pin: tests/test_oracle_vs_ref.py checks every function here against the compiled reference
// (oracle/_ref/libref_oracle.so) and against the reference's own known-answer tests
// (tests/golden/known_answers.json), see DESIGN.md "Oracle".
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <limits>
#include <queue>
#include <random>
#include <string>
#include <thread>
#include <unordered_map>
#include <unordered_set>
#include <vector>

namespace {

constexpr uint32_t kEmpty = 0xFFFFFFFFu;  // include/index/graph/graph.hpp:49 (kEmptyId = -1)

enum Metric { kL2 = 0, kIP = 1, kCOS = 2 };  // include/utils/metric_type.hpp:27-32
enum Quant { kNone = 0, kSQ8 = 1, kSQ4 = 2 };

// ---- distance kernels: the "generic" variants, sequential accumulation -------------------------
// include/simd/distance_l2.ipp:41-50
float l2_sqr(const float *x, const float *y, size_t dim) {
  float sum = 0.0F;
  for (size_t i = 0; i < dim; ++i) {
    float diff = x[i] - y[i];
    sum += diff * diff;
  }
  return sum;
}
// include/simd/distance_ip.ipp:45-52
float ip_sqr(const float *x, const float *y, size_t dim) {
  float sum = 0.0F;
  for (size_t i = 0; i < dim; ++i) sum += x[i] * y[i];
  return -sum;
}
// include/simd/distance_l2.ipp:187-201
float l2_sqr_sq8(const uint8_t *x, const uint8_t *y, size_t dim, const float *mn, const float *mx) {
  const float inv = 1.0F / 255.0F;
  float sum = 0.0F;
  for (size_t i = 0; i < dim; ++i) {
    float scale = (mx[i] - mn[i]) * inv;
    float diff = static_cast<float>(x[i]) - static_cast<float>(y[i]);
    float sd = diff * scale;
    sum += sd * sd;
  }
  return sum;
}
// include/simd/distance_ip.ipp:177-191
float ip_sqr_sq8(const uint8_t *x, const uint8_t *y, size_t dim, const float *mn, const float *mx) {
  const float inv = 1.0F / 255.0F;
  float sum = 0.0F;
  for (size_t i = 0; i < dim; ++i) {
    float scale = (mx[i] - mn[i]) * inv;
    float xv = mn[i] + static_cast<float>(x[i]) * scale;
    float yv = mn[i] + static_cast<float>(y[i]) * scale;
    sum += xv * yv;
  }
  return -sum;
}
// include/simd/distance_l2.ipp:209-237. NOTE the nibble pairing: the LOW nibble is scored with
// (min,max)[i] and the HIGH nibble with (min,max)[i+1], while the encoder (sq4.hpp:139-149) puts
// dim i in the HIGH nibble -- reproduced on purpose.
float l2_sqr_sq4(const uint8_t *x, const uint8_t *y, size_t dim, const float *mn, const float *mx) {
  const float inv = 1.0F / 15.0F;
  float sum = 0.0F;
  size_t b = 0;
  for (size_t i = 0; i < dim; i += 2, ++b) {
    float s_lo = (mx[i] - mn[i]) * inv;
    float d_lo = static_cast<float>(x[b] & 0x0F) - static_cast<float>(y[b] & 0x0F);
    float sd_lo = d_lo * s_lo;
    sum += sd_lo * sd_lo;
    if (i + 1 < dim) {
      float s_hi = (mx[i + 1] - mn[i + 1]) * inv;
      float d_hi = static_cast<float>((x[b] >> 4) & 0x0F) - static_cast<float>((y[b] >> 4) & 0x0F);
      float sd_hi = d_hi * s_hi;
      sum += sd_hi * sd_hi;
    }
  }
  return sum;
}
// include/simd/distance_ip.ipp:380-407 (same nibble pairing)
float ip_sqr_sq4(const uint8_t *x, const uint8_t *y, size_t dim, const float *mn, const float *mx) {
  const float inv = 1.0F / 15.0F;
  float sum = 0.0F;
  size_t b = 0;
  for (size_t i = 0; i < dim; i += 2, ++b) {
    float s_lo = (mx[i] - mn[i]) * inv;
    float xv = mn[i] + static_cast<float>(x[b] & 0x0F) * s_lo;
    float yv = mn[i] + static_cast<float>(y[b] & 0x0F) * s_lo;
    sum += xv * yv;
    if (i + 1 < dim) {
      float s_hi = (mx[i + 1] - mn[i + 1]) * inv;
      float xh = mn[i + 1] + static_cast<float>((x[b] >> 4) & 0x0F) * s_hi;
      float yh = mn[i + 1] + static_cast<float>((y[b] >> 4) & 0x0F) * s_hi;
      sum += xh * yh;
    }
  }
  return -sum;
}

// include/utils/data_utils.hpp:36-46: float accumulation, 1.0/sqrtf in double, rounded to float.
void normalize(float *v, size_t dim) {
  float sum = 0;
  for (size_t i = 0; i < dim; ++i) sum += v[i] * v[i];
  sum = static_cast<float>(1.0 / std::sqrt(sum));
  for (size_t i = 0; i < dim; ++i) v[i] *= sum;
}

// ---- scalar quantizers -------------------------------------------------------------------------
// include/space/quant/sq8.hpp:120-132 and sq4.hpp:120-132 (levels = 255 or 15)
uint8_t quantize(float v, float mn, float mx, int levels) {
  if (mx == mn) return 0;
  if (v >= mx) return static_cast<uint8_t>(levels);
  if (v <= mn) return 0;
  float scaled = (v - mn) / (mx - mn);
  return static_cast<uint8_t>(scaled * static_cast<float>(levels));
}
// sq8.hpp:78-111 (init max=lowest, min=max(); strict < / > updates)
void fit_minmax(const float *data, size_t n, uint32_t dim, float *mn, float *mx) {
  for (uint32_t d = 0; d < dim; d++) {
    mn[d] = std::numeric_limits<float>::max();
    mx[d] = std::numeric_limits<float>::lowest();
  }
  for (size_t i = 0; i < n; i++)
    for (uint32_t d = 0; d < dim; d++) {
      float v = data[i * dim + d];
      if (v < mn[d]) mn[d] = v;
      if (v > mx[d]) mx[d] = v;
    }
}
// sq8.hpp:139-143
void encode_sq8(const float *raw, uint8_t *out, uint32_t dim, const float *mn, const float *mx) {
  for (uint32_t i = 0; i < dim; i++) out[i] = quantize(raw[i], mn[i], mx[i], 255);
}
// sq4.hpp:139-149: dim i in the high nibble, dim i+1 in the low nibble, odd tail -> low nibble 0
void encode_sq4(const float *raw, uint8_t *out, uint32_t dim, const float *mn, const float *mx) {
  for (uint32_t i = 0; i < dim; i += 2) {
    uint8_t first = quantize(raw[i], mn[i], mx[i], 15);
    uint8_t second = 0;
    if (i + 1 < dim) second = quantize(raw[i + 1], mn[i + 1], mx[i + 1], 15);
    out[i / 2] = static_cast<uint8_t>((first << 4) | second);
  }
}

// ---- LinearPool + visited bitmap ---------------------------------------------------------------
// include/utils/query_utils.hpp:237-312. Sorted ascending; bit 31 of the id = "expanded".
struct Pool {
  struct Item {
    uint32_t id;
    float dist;
  };
  size_t size = 0, cur = 0, cap;
  std::vector<Item> data;
  std::vector<uint64_t> vis;  // query_utils.hpp:69-115 DynamicBitset
  Pool(size_t n, size_t capacity) : cap(capacity), data(capacity + 1, Item{0, 0.f}), vis((n + 63) / 64, 0) {}
  bool get(size_t p) const { return (vis[p >> 6] >> (p & 63)) & 1ULL; }
  void set(size_t p) { vis[p >> 6] |= 1ULL << (p & 63); }
  // :240-252 upper bound: first index whose distance is > dist (ties go after equal entries)
  int find(float dist) const {
    int l = 0, r = static_cast<int>(size);
    while (l < r) {
      int mid = (l + r) / 2;
      if (data[mid].dist > dist) r = mid; else l = mid + 1;
    }
    return l;
  }
  // :254-272
  bool insert(uint32_t u, float dist) {
    if (size == cap && dist >= data[size - 1].dist) return false;
    int lo = find(dist);
    std::memmove(&data[lo + 1], &data[lo], (size - lo) * sizeof(Item));
    data[lo] = Item{u, dist};
    if (size < cap) size++;
    if (static_cast<size_t>(lo) < cur) cur = lo;
    return true;
  }
  // :284-293
  uint32_t pop() {
    data[cur].id |= 1u << 31;
    size_t pre = cur;
    while (cur < size && (data[cur].id >> 31) != 0) cur++;
    return data[pre].id & 0x7FFFFFFFu;
  }
  bool has_next() const { return cur < size; }
  uint32_t id(size_t i) const { return data[i].id & 0x7FFFFFFFu; }
};

// ---- the index ---------------------------------------------------------------------------------
struct Index {
  int metric = kL2, quant = kNone;
  uint32_t dim = 0, capacity = 0, n = 0, R = 32;
  std::vector<float> raw;          // [n, dim]  (RawSpace rows; normalised in place for COS)
  std::vector<uint8_t> valid;      // SequentialStorage validity bit per row (raw space)
  std::vector<uint8_t> codes;      // [n, row_bytes]
  std::vector<float> mn, mx;       // quantizer
  std::vector<uint32_t> adj;       // [capacity or n, R], kEmpty padded
  std::vector<uint32_t> levels;    // overlay levels per node
  std::vector<std::vector<uint32_t>> lists;  // overlay lists: levels*R ids per node
  uint32_t ep = 0;
  bool has_overlay = false;
  // update-job context (include/executor/jobs/job_context.hpp:25-29)
  std::unordered_map<uint32_t, std::vector<uint32_t>> inserted_edges;
  std::unordered_set<uint32_t> removed_vertices;
  std::unordered_map<uint32_t, std::vector<uint32_t>> removed_node_nbrs;
  std::unordered_set<uint32_t> graph_removed;  // rows whose Graph storage validity bit was cleared

  size_t row_bytes() const { return quant == kSQ8 ? dim : (dim + 1) / 2; }

  // RawSpace::get_distance, raw_space.hpp:178-180 (+ metric function :117-129)
  float raw_dist(const float *a, const float *b) const {
    return metric == kL2 ? l2_sqr(a, b, dim) : ip_sqr(a, b, dim);
  }
  float raw_dist_id(uint32_t i, uint32_t j) const { return raw_dist(&raw[size_t(i) * dim], &raw[size_t(j) * dim]); }
  // SQ8Space / SQ4Space metric functions: sq8_space.hpp:91-104, sq4_space.hpp:94-107
  float code_dist(const uint8_t *a, const uint8_t *b) const {
    if (quant == kSQ8)
      return metric == kL2 ? l2_sqr_sq8(a, b, dim, mn.data(), mx.data()) : ip_sqr_sq8(a, b, dim, mn.data(), mx.data());
    return metric == kL2 ? l2_sqr_sq4(a, b, dim, mn.data(), mx.data()) : ip_sqr_sq4(a, b, dim, mn.data(), mx.data());
  }
};

// QueryComputer of the *search* space: raw_space.hpp:255-305, sq8_space.hpp:255-298,
// sq4_space.hpp:256-300. Raw + COS normalises the query; SQ spaces encode it un-normalised.
struct QueryComputer {
  const Index &ix;
  std::vector<float> q;
  std::vector<uint8_t> qc;
  bool use_codes;
  QueryComputer(const Index &index, const float *query, bool search_space) : ix(index) {
    use_codes = search_space && ix.quant != kNone;
    if (!use_codes) {
      q.assign(query, query + ix.dim);
      if (ix.metric == kCOS) normalize(q.data(), ix.dim);
    } else {
      qc.assign(ix.row_bytes(), 0);
      if (ix.quant == kSQ8) encode_sq8(query, qc.data(), ix.dim, ix.mn.data(), ix.mx.data());
      else encode_sq4(query, qc.data(), ix.dim, ix.mn.data(), ix.mx.data());
    }
  }
  // by-id constructor (raw_space.hpp:276-281 / sq8_space.hpp:271-276): copies the stored row
  QueryComputer(const Index &index, uint32_t id, bool search_space) : ix(index) {
    use_codes = search_space && ix.quant != kNone;
    if (!use_codes) q.assign(&ix.raw[size_t(id) * ix.dim], &ix.raw[size_t(id + 1) * ix.dim]);
    else qc.assign(&ix.codes[size_t(id) * ix.row_bytes()], &ix.codes[size_t(id + 1) * ix.row_bytes()]);
  }
  float operator()(uint32_t u) const {
    if (!use_codes) {
      if (!ix.valid[u]) return std::numeric_limits<float>::max();  // raw_space.hpp:298-300
      return ix.raw_dist(q.data(), &ix.raw[size_t(u) * ix.dim]);
    }
    return ix.code_dist(qc.data(), &ix.codes[size_t(u) * ix.row_bytes()]);
  }
};

// include/index/graph/overlay_graph.hpp:122-144 (the scan keeps walking the OLD list after u moves)
void overlay_initialize(const Index &ix, Pool &pool, const QueryComputer &qc, uint64_t *ov_calls) {
  uint32_t u = ix.ep;
  float cur = qc(u);
  uint64_t calls = 1;
  for (int level = static_cast<int>(ix.levels[u]); level > 0; --level) {
    bool changed = true;
    while (changed) {
      changed = false;
      const uint32_t *list = ix.lists[u].data() + size_t(level - 1) * ix.R;
      for (uint32_t i = 0; i < ix.R && list[i] != kEmpty; ++i) {
        uint32_t v = list[i];
        float d = qc(v);
        calls++;
        if (d < cur) {
          cur = d;
          u = v;
          changed = true;
        }
      }
    }
  }
  pool.insert(u, cur);
  pool.set(u);
  if (ov_calls) *ov_calls = calls;
}

// include/executor/jobs/graph_search_job.hpp:302-335 (search_solo) == :221-258 (coroutine search)
// graph.hpp:148-158 initialize_search. `updated` selects search_solo_updated (:373-417).
void traverse(const Index &ix, const QueryComputer &qc, Pool &pool, uint64_t *stats, bool updated = false) {
  uint64_t ndc = 0, hops = 0, ov = 0;
  if (ix.has_overlay) {
    overlay_initialize(ix, pool, qc, &ov);
  } else {
    pool.insert(ix.ep, qc(ix.ep));
    pool.set(ix.ep);
  }
  while (pool.has_next()) {
    uint32_t u = pool.pop();
    hops++;
    if (updated && ix.removed_node_nbrs.count(u)) {
      for (uint32_t w : ix.removed_node_nbrs.at(u)) {
        if (pool.get(w)) continue;
        pool.set(w);
        pool.insert(w, qc(w));
      }
      continue;
    }
    const uint32_t *row = &ix.adj[size_t(u) * ix.R];
    for (uint32_t i = 0; i < ix.R; ++i) {
      uint32_t v = row[i];
      if (v == kEmpty) break;
      if (pool.get(v)) continue;
      pool.set(v);
      ndc++;
      pool.insert(v, qc(v));
    }
  }
  if (stats) {
    stats[0] = ndc;
    stats[1] = hops;
    stats[2] = ov;
  }
}

// python/include/index.hpp:450-488. `src` has ef entries of which only the first topk were written
// by search (graph_search_job.hpp:254-256); the rest keep their zero initialisation
// (index.hpp:239 / :301) and are re-ranked as id 0 -- reproduced.
void rerank(const Index &ix, const std::vector<uint32_t> &src, const float *query, uint32_t ef,
            uint32_t topk, uint32_t *ids, float *dist) {
  QueryComputer qc(ix, query, /*search_space=*/false);
  std::priority_queue<std::pair<float, uint32_t>, std::vector<std::pair<float, uint32_t>>, std::greater<>> pq;
  for (size_t i = 0; i < ef; i++) pq.push({qc(src[i]), src[i]});
  for (size_t i = 0; i < topk; i++) {
    ids[i] = pq.top().second;
    if (dist) dist[i] = pq.top().first;
    pq.pop();
  }
}

// PyIndex::search / one query of PyIndex::batch_search: index.hpp:236-258, :289-346
void search_one(const Index &ix, const float *query, uint32_t topk, uint32_t ef, uint32_t *ids,
                float *dist, bool with_distance_api) {
  std::vector<uint32_t> res_pool(ef, 0);
  {
    QueryComputer qc(ix, query, /*search_space=*/true);
    Pool pool(ix.n, ef);
    traverse(ix, qc, pool, nullptr);
    for (uint32_t i = 0; i < topk; i++) {
      res_pool[i] = pool.id(i);
      if (dist && ix.quant == kNone) dist[i] = pool.data[i].dist;
    }
  }
  if (ix.quant == kNone || with_distance_api) {
    // batch_search_with_distance on Linux returns the un-reranked ids for SQ spaces (index.hpp:395-418)
    std::copy(res_pool.begin(), res_pool.begin() + topk, ids);
  } else {
    rerank(ix, res_pool, query, ef, topk, ids, nullptr);
  }
}

// ---- HNSW build --------------------------------------------------------------------------------
// include/index/graph/hnsw/hnswlib.hpp (HNSWImpl) + hnsw_builder.hpp:98-194, single-threaded.
struct Hnsw {
  using Pair = std::pair<float, uint32_t>;
  struct ByFirst {  // hnswlib.hpp:138-143
    bool operator()(const Pair &a, const Pair &b) const { return a.first < b.first; }
  };
  using MaxHeap = std::priority_queue<Pair, std::vector<Pair>, ByFirst>;

  const Index &ix;
  size_t M, M0, efc;
  double mult;
  int maxlevel = -1;
  uint32_t enterpoint = kEmpty;
  std::vector<int> element_levels;
  std::vector<std::vector<uint32_t>> l0;                  // level-0 link lists (count implicit)
  std::vector<std::vector<std::vector<uint32_t>>> upper;  // [node][level-1] link lists
  std::default_random_engine level_generator;             // hnswlib.hpp:85, seeded 100 (:91,:103)
  std::vector<uint16_t> visited;                          // visited_list_pool.hpp tags
  uint16_t tag = 0;
  uint64_t dist_calls = 0;

  Hnsw(const Index &index, size_t max_edge, size_t ef_construction)
      : ix(index), M(max_edge), M0(max_edge * 2), efc(std::max(ef_construction, max_edge)),
        mult(1 / std::log(1.0 * static_cast<double>(max_edge))), element_levels(index.n, 0),
        l0(index.n), upper(index.n), visited(index.n, 0) {
    level_generator.seed(100);
  }
  float dist(uint32_t a, uint32_t b) {
    dist_calls++;
    return ix.raw_dist_id(a, b);
  }
  // hnswlib.hpp:182-186
  size_t random_level() {
    std::uniform_real_distribution<double> distribution(0.0, 1.0);
    double r = -std::log(distribution(level_generator)) * mult;
    return static_cast<size_t>(r);
  }
  std::vector<uint32_t> &links(uint32_t id, int level) { return level == 0 ? l0[id] : upper[id][level - 1]; }

  // hnswlib.hpp:291-354
  void heuristic(MaxHeap &top, size_t m) {
    if (top.size() < m) return;
    std::priority_queue<Pair> closest;
    std::vector<Pair> ret;
    while (!top.empty()) {
      closest.emplace(-top.top().first, top.top().second);
      top.pop();
    }
    while (!closest.empty()) {
      if (ret.size() >= m) break;
      Pair cur = closest.top();
      float d_q = -cur.first;
      closest.pop();
      bool good = true;
      for (const Pair &s : ret) {
        if (dist(s.second, cur.second) < d_q) {
          good = false;
          break;
        }
      }
      if (good) ret.push_back(cur);
    }
    for (const Pair &p : ret) top.emplace(-p.first, p.second);
  }

  // hnswlib.hpp:373-489
  MaxHeap search_layer(uint32_t ep_id, uint32_t label, int layer) {
    if (++tag == 0) {  // visited_list_pool.hpp: reset on wrap
      std::fill(visited.begin(), visited.end(), 0);
      tag = 1;
    }
    MaxHeap top, cand;
    float d = dist(label, ep_id);
    top.emplace(d, ep_id);
    float lower = d;
    cand.emplace(-d, ep_id);
    visited[ep_id] = tag;
    while (!cand.empty()) {
      Pair cur = cand.top();
      if ((-cur.first) > lower && top.size() == efc) break;
      cand.pop();
      const std::vector<uint32_t> &lst = links(cur.second, layer);
      for (uint32_t c : lst) {
        if (visited[c] == tag) continue;
        visited[c] = tag;
        float d1 = dist(label, c);
        if (top.size() < efc || lower > d1) {
          cand.emplace(-d1, c);
          top.emplace(d1, c);
          if (top.size() > efc) top.pop();
          if (!top.empty()) lower = top.top().first;
        }
      }
    }
    return top;
  }

  // hnswlib.hpp:512-635 (isUpdate == false)
  uint32_t connect(uint32_t cur_c, MaxHeap &top, int level) {
    size_t mcurmax = level != 0 ? M : M0;
    heuristic(top, M);
    std::vector<uint32_t> selected;
    selected.reserve(M);
    while (!top.empty()) {
      selected.push_back(top.top().second);
      top.pop();
    }
    uint32_t next_ep = selected.back();
    links(cur_c, level) = selected;
    for (uint32_t nb : selected) {
      std::vector<uint32_t> &other = links(nb, level);
      if (other.size() < mcurmax) {
        other.push_back(cur_c);
      } else {
        float d_max = dist(cur_c, nb);
        MaxHeap cands;
        cands.emplace(d_max, cur_c);
        for (uint32_t o : other) cands.emplace(dist(o, nb), o);
        heuristic(cands, mcurmax);
        other.clear();
        while (!cands.empty()) {
          other.push_back(cands.top().second);
          cands.pop();
        }
      }
    }
    return next_ep;
  }

  // hnswlib.hpp:652-751 (labels == internal ids when inserted in order from one thread)
  void add_point(uint32_t id) {
    int cur_level = static_cast<int>(random_level());
    element_levels[id] = cur_level;
    int maxlevel_copy = maxlevel;
    uint32_t curr = enterpoint;
    if (cur_level != 0) upper[id].assign(cur_level, {});
    if (curr != kEmpty) {
      if (cur_level < maxlevel_copy) {
        float curdist = dist(id, curr);
        for (int level = maxlevel_copy; level > cur_level; level--) {
          bool changed = true;
          while (changed) {
            changed = false;
            // NB: hnswlib re-reads the list of the *current* node each while-iteration but keeps
            // scanning the list it started with inside the for loop.
            const std::vector<uint32_t> lst = links(curr, level);
            for (uint32_t cand : lst) {
              float d = dist(id, cand);
              if (d < curdist) {
                curdist = d;
                curr = cand;
                changed = true;
              }
            }
          }
        }
      }
      for (int level = std::min(cur_level, maxlevel_copy); level >= 0; level--) {
        MaxHeap top = search_layer(curr, id, level);
        curr = connect(id, top, level);
      }
    } else {
      enterpoint = 0;
      maxlevel = cur_level;
    }
    if (cur_level > maxlevel_copy) {
      enterpoint = id;
      maxlevel = cur_level;
    }
  }
};

// hnsw_builder.hpp:98-194: export level-0 lists (-1 padded) and overlay lists (levels*R, R/2 used)
void build_hnsw(Index &ix, uint32_t efc, uint64_t *dist_calls) {
  Hnsw h(ix, ix.R / 2, efc);
  for (uint32_t i = 0; i < ix.n; i++) h.add_point(i);
  ix.adj.assign(size_t(std::max(ix.capacity, ix.n)) * ix.R, kEmpty);
  ix.levels.assign(ix.n, 0);
  ix.lists.assign(ix.n, {});
  for (uint32_t i = 0; i < ix.n; i++) {
    std::copy(h.l0[i].begin(), h.l0[i].end(), &ix.adj[size_t(i) * ix.R]);
    int level = h.element_levels[i];
    ix.levels[i] = level;
    if (level > 0) {
      ix.lists[i].assign(size_t(level) * ix.R, kEmpty);
      for (int j = 1; j <= level; j++)
        std::copy(h.upper[i][j - 1].begin(), h.upper[i][j - 1].end(), &ix.lists[i][size_t(j - 1) * ix.R]);
    }
  }
  ix.ep = h.enterpoint;
  ix.has_overlay = true;
  if (dist_calls) *dist_calls = h.dist_calls;
}

// PyIndex::fit spaces part: index.hpp:197-206; RawSpace::fit raw_space.hpp:131-139 (COS normalises
// the rows before storing them, and the SQ space is then fitted on the normalised rows because the
// reference normalises the caller's buffer in place); SQ8Space::fit sq8_space.hpp:116-127.
void fit_spaces(Index &ix, const float *data, uint32_t n) {
  ix.n = n;
  ix.raw.assign(data, data + size_t(n) * ix.dim);
  ix.valid.assign(std::max(ix.capacity, n), 0);
  std::fill(ix.valid.begin(), ix.valid.begin() + n, 1);
  if (ix.metric == kCOS)
    for (uint32_t i = 0; i < n; i++) normalize(&ix.raw[size_t(i) * ix.dim], ix.dim);
  if (ix.quant != kNone) {
    ix.mn.assign(ix.dim, 0.f);
    ix.mx.assign(ix.dim, 0.f);
    fit_minmax(ix.raw.data(), n, ix.dim, ix.mn.data(), ix.mx.data());
    size_t rb = ix.row_bytes();
    ix.codes.assign(size_t(n) * rb, 0);
    for (uint32_t i = 0; i < n; i++) {
      if (ix.quant == kSQ8) encode_sq8(&ix.raw[size_t(i) * ix.dim], &ix.codes[i * rb], ix.dim, ix.mn.data(), ix.mx.data());
      else encode_sq4(&ix.raw[size_t(i) * ix.dim], &ix.codes[i * rb], ix.dim, ix.mn.data(), ix.mx.data());
    }
  }
}

// ---- incremental update: include/executor/jobs/graph_update_job.hpp ----------------------------
// :105-137 update(node): candidates = current edges (+ 2-hop of removed nbrs) + inserted reverse
// edges; keep the R closest by the search-space distance; unfilled slots are written as 0 (:132-136
// value-initialised vector), not -1.
void update_node(Index &ix, uint32_t node) {
  std::unordered_set<uint32_t> cand;
  const uint32_t *cur = &ix.adj[size_t(node) * ix.R];
  for (uint32_t i = 0; i < ix.R; i++) {
    uint32_t nb = cur[i];
    if (nb == kEmpty) break;
    if (ix.removed_vertices.count(nb))
      for (uint32_t s : ix.removed_node_nbrs.at(nb)) cand.insert(s);
    cand.insert(nb);
  }
  if (ix.inserted_edges.count(node))
    for (uint32_t s : ix.inserted_edges.at(node)) cand.insert(s);
  QueryComputer qc(ix, node, /*search_space=*/true);
  Pool pool(ix.n, ix.R);
  for (uint32_t nb : cand) pool.insert(nb, qc(nb));
  std::vector<uint32_t> upd(ix.R, 0);
  for (uint32_t i = 0; i < ix.R && i < pool.size; i++) upd[i] = pool.id(i);
  // Graph::update -> SequentialStorage::update only writes when the row is still valid
  // (sequential_storage.hpp:98-104), i.e. the node has not been removed.
  if (ix.graph_removed.count(node)) return;
  std::copy(upd.begin(), upd.end(), &ix.adj[size_t(node) * ix.R]);
}

}  // namespace

// ================================== C API ========================================================
extern "C" {

void *orc_index_create(int metric, int quant, uint32_t capacity, uint32_t dim) {
  auto *ix = new Index();
  ix->metric = metric;
  ix->quant = quant;
  ix->capacity = capacity;
  ix->dim = dim;
  return ix;
}
void orc_index_destroy(void *h) { delete static_cast<Index *>(h); }

int orc_index_fit(void *h, const float *data, uint32_t n, uint32_t efc, uint64_t *dist_calls) {
  auto &ix = *static_cast<Index *>(h);
  if (ix.quant != kNone && n > ix.capacity) return -1;  // sq8_space.hpp:117-119 throws
  fit_spaces(ix, data, n);
  build_hnsw(ix, efc, dist_calls);
  return 0;
}
// spaces only (no graph): lets tests hand in a graph built elsewhere
int orc_index_fit_spaces(void *h, const float *data, uint32_t n) {
  fit_spaces(*static_cast<Index *>(h), data, n);
  return 0;
}
// adjacency [n, R]; levels [n]; overlay lists concatenated in node order (levels[i]*R ids each)
int orc_index_set_graph(void *h, const uint32_t *adj, const uint32_t *levels, const uint32_t *lists_flat,
                        uint32_t ep, int has_overlay) {
  auto &ix = *static_cast<Index *>(h);
  ix.adj.assign(size_t(std::max(ix.capacity, ix.n)) * ix.R, kEmpty);
  std::copy(adj, adj + size_t(ix.n) * ix.R, ix.adj.begin());
  ix.has_overlay = has_overlay != 0;
  ix.ep = ep;
  ix.levels.assign(ix.n, 0);
  ix.lists.assign(ix.n, {});
  if (ix.has_overlay) {
    size_t off = 0;
    for (uint32_t i = 0; i < ix.n; i++) {
      ix.levels[i] = levels[i];
      size_t cnt = size_t(levels[i]) * ix.R;
      ix.lists[i].assign(lists_flat + off, lists_flat + off + cnt);
      off += cnt;
    }
  }
  return 0;
}
int orc_index_graph_info(void *h, uint32_t *info) {
  auto &ix = *static_cast<Index *>(h);
  info[0] = std::max(ix.capacity, ix.n);
  info[1] = ix.R;
  info[2] = ix.ep;
  info[3] = ix.has_overlay;
  info[4] = ix.n;
  return 0;
}
int orc_index_graph_rows(void *h, uint32_t first, uint32_t count, uint32_t *out) {
  auto &ix = *static_cast<Index *>(h);
  std::memcpy(out, &ix.adj[size_t(first) * ix.R], size_t(count) * ix.R * 4);
  return 0;
}
int orc_index_overlay_levels(void *h, uint32_t count, uint32_t *levels) {
  auto &ix = *static_cast<Index *>(h);
  std::copy(ix.levels.begin(), ix.levels.begin() + count, levels);
  return 0;
}
int orc_index_overlay_list(void *h, uint32_t id, uint32_t *out) {
  auto &l = static_cast<Index *>(h)->lists[id];
  std::copy(l.begin(), l.end(), out);
  return static_cast<int>(l.size());
}
int orc_index_raw_row(void *h, uint32_t id, float *out) {
  auto &ix = *static_cast<Index *>(h);
  std::memcpy(out, &ix.raw[size_t(id) * ix.dim], ix.dim * 4);
  return 0;
}
int orc_index_code_row(void *h, uint32_t id, uint8_t *out, uint32_t nbytes) {
  auto &ix = *static_cast<Index *>(h);
  if (ix.quant == kNone) return -1;
  std::memcpy(out, &ix.codes[size_t(id) * ix.row_bytes()], nbytes);
  return 0;
}
int orc_index_quant_minmax(void *h, float *mn, float *mx) {
  auto &ix = *static_cast<Index *>(h);
  if (ix.quant == kNone) return -1;
  std::copy(ix.mn.begin(), ix.mn.end(), mn);
  std::copy(ix.mx.begin(), ix.mx.end(), mx);
  return 0;
}
float orc_index_space_distance(void *h, uint32_t i, uint32_t j) {
  auto &ix = *static_cast<Index *>(h);
  if (ix.quant == kNone) return ix.raw_dist_id(i, j);
  size_t rb = ix.row_bytes();
  return ix.code_dist(&ix.codes[i * rb], &ix.codes[j * rb]);
}

int orc_index_search(void *h, const float *q, uint32_t topk, uint32_t ef, uint32_t *ids) {
  search_one(*static_cast<Index *>(h), q, topk, ef, ids, nullptr, false);
  return 0;
}
// threads > 1 splits queries round-robin; every query is an independent sequential traversal, so
// results do not depend on the thread count (SURVEY.md 8c "Determinism").
int orc_index_batch_search(void *h, const float *q, uint32_t nq, uint32_t topk, uint32_t ef,
                           uint32_t threads, uint32_t *ids) {
  auto &ix = *static_cast<Index *>(h);
  if (threads < 1) threads = 1;
  std::vector<std::thread> th;
  for (uint32_t t = 0; t < threads; t++)
    th.emplace_back([&, t] {
      for (uint32_t i = t; i < nq; i += threads)
        search_one(ix, q + size_t(i) * ix.dim, topk, ef, ids + size_t(i) * topk, nullptr, false);
    });
  for (auto &x : th) x.join();
  return 0;
}
int orc_index_batch_search_with_distance(void *h, const float *q, uint32_t nq, uint32_t topk,
                                         uint32_t ef, uint32_t threads, uint32_t *ids, float *dist,
                                         int *has_dist) {
  auto &ix = *static_cast<Index *>(h);
  (void)threads;
  *has_dist = ix.quant == kNone ? 1 : 0;
  for (uint32_t i = 0; i < nq; i++)
    search_one(ix, q + size_t(i) * ix.dim, topk, ef, ids + size_t(i) * topk,
               ix.quant == kNone ? dist + size_t(i) * topk : nullptr, true);
  return 0;
}
// whole pool + counters: stats = {level-0 distance calls, hops, overlay distance calls}
int orc_index_traverse(void *h, const float *q, uint32_t ef, uint32_t *pool_ids, float *pool_dist,
                       uint32_t *pool_size, uint64_t *stats) {
  auto &ix = *static_cast<Index *>(h);
  QueryComputer qc(ix, q, true);
  Pool pool(ix.n, ef);
  traverse(ix, qc, pool, stats);
  *pool_size = static_cast<uint32_t>(pool.size);
  for (size_t i = 0; i < pool.size; i++) {
    pool_ids[i] = pool.id(i);
    pool_dist[i] = pool.data[i].dist;
  }
  return 0;
}

// GraphUpdateJob::insert_and_update, graph_update_job.hpp:66-89. The new row goes into the SEARCH
// space only (:75); the raw space used for rerank is not extended (reference quirk) unless the two
// are the same object (quant none).
int64_t orc_index_insert(void *h, const float *v, uint32_t ef) {
  auto &ix = *static_cast<Index *>(h);
  std::vector<uint32_t> res(ix.R, kEmpty);
  {
    QueryComputer qc(ix, v, true);
    Pool pool(ix.n, ef);
    traverse(ix, qc, pool, nullptr);
    for (uint32_t i = 0; i < ix.R; i++) res[i] = pool.id(i);  // k = max_nbrs; zeros past pool.size
  }
  uint32_t cap = std::max(ix.capacity, 0u);
  if (ix.n >= cap) return -1;
  uint32_t id = ix.n;
  std::copy(res.begin(), res.end(), &ix.adj[size_t(id) * ix.R]);
  if (ix.quant == kNone) {
    ix.raw.insert(ix.raw.end(), v, v + ix.dim);
    if (ix.metric == kCOS) normalize(&ix.raw[size_t(id) * ix.dim], ix.dim);
    ix.valid[id] = 1;
  } else {
    size_t rb = ix.row_bytes();
    ix.codes.resize(size_t(id + 1) * rb);
    if (ix.quant == kSQ8) encode_sq8(v, &ix.codes[id * rb], ix.dim, ix.mn.data(), ix.mx.data());
    else encode_sq4(v, &ix.codes[id * rb], ix.dim, ix.mn.data(), ix.mx.data());
  }
  ix.n++;
  if (ix.has_overlay) {
    ix.levels.push_back(0);
    ix.lists.emplace_back();
  }
  for (uint32_t i = 0; i < ix.R; i++)
    if (res[i] != kEmpty) ix.inserted_edges[res[i]].push_back(id);
  std::vector<uint32_t> keys;
  for (auto &kv : ix.inserted_edges) keys.push_back(kv.first);
  for (uint32_t k : keys) update_node(ix, k);
  ix.inserted_edges.clear();
  return id;
}
// graph_update_job.hpp:91-103
int orc_index_remove(void *h, uint32_t id) {
  auto &ix = *static_cast<Index *>(h);
  const uint32_t *row = &ix.adj[size_t(id) * ix.R];
  for (uint32_t i = 0; i < ix.R; i++) {
    if (row[i] == kEmpty) break;
    ix.removed_node_nbrs[id].push_back(row[i]);
  }
  ix.removed_vertices.insert(id);
  ix.graph_removed.insert(id);
  if (ix.quant == kNone) ix.valid[id] = 0;
  return 0;
}

// kernels
float orc_l2_sqr(const float *x, const float *y, size_t d) { return l2_sqr(x, y, d); }
float orc_ip_sqr(const float *x, const float *y, size_t d) { return ip_sqr(x, y, d); }
float orc_l2_sqr_sq8(const uint8_t *x, const uint8_t *y, size_t d, const float *mn, const float *mx) { return l2_sqr_sq8(x, y, d, mn, mx); }
float orc_ip_sqr_sq8(const uint8_t *x, const uint8_t *y, size_t d, const float *mn, const float *mx) { return ip_sqr_sq8(x, y, d, mn, mx); }
float orc_l2_sqr_sq4(const uint8_t *x, const uint8_t *y, size_t d, const float *mn, const float *mx) { return l2_sqr_sq4(x, y, d, mn, mx); }
float orc_ip_sqr_sq4(const uint8_t *x, const uint8_t *y, size_t d, const float *mn, const float *mx) { return ip_sqr_sq4(x, y, d, mn, mx); }
void orc_normalize(float *v, size_t d) { normalize(v, d); }

int orc_sq8_fit_encode(const float *data, size_t n, uint32_t dim, float *mn, float *mx, uint8_t *codes) {
  fit_minmax(data, n, dim, mn, mx);
  for (size_t i = 0; i < n; i++) encode_sq8(data + i * dim, codes + i * dim, dim, mn, mx);
  return 0;
}
int orc_sq4_fit_encode(const float *data, size_t n, uint32_t dim, float *mn, float *mx, uint8_t *codes) {
  fit_minmax(data, n, dim, mn, mx);
  size_t rb = (dim + 1) / 2;
  for (size_t i = 0; i < n; i++) encode_sq4(data + i * dim, codes + i * rb, dim, mn, mx);
  return 0;
}
uint8_t orc_sq8_quantize(float v, float mn, float mx) { return quantize(v, mn, mx, 255); }
uint8_t orc_sq4_quantize(float v, float mn, float mx) { return quantize(v, mn, mx, 15); }

// include/utils/evaluate.hpp:30-62: fp32 L2 against every row, std::sort by distance only
int orc_find_exact_gt(const float *q, uint32_t nq, const float *data, uint32_t n, uint32_t dim,
                      uint32_t topk, uint32_t threads, uint32_t *ids) {
  if (threads < 1) threads = 1;
  std::vector<std::thread> th;
  for (uint32_t t = 0; t < threads; t++)
    th.emplace_back([&, t] {
      std::vector<std::pair<uint32_t, float>> d(n);
      for (uint32_t i = t; i < nq; i += threads) {
        for (uint32_t j = 0; j < n; j++) d[j] = {j, l2_sqr(q + size_t(i) * dim, data + size_t(j) * dim, dim)};
        std::sort(d.begin(), d.end(), [](const auto &a, const auto &b) { return a.second < b.second; });
        for (uint32_t j = 0; j < topk; j++) ids[size_t(i) * topk + j] = d[j].first;
      }
    });
  for (auto &x : th) x.join();
  return 0;
}
// exact top-k in (distance, id) order with distances; the documented tie rule of the GPU path
int orc_bruteforce_topk(int metric, const float *q, uint32_t nq, const float *data, uint32_t n,
                        uint32_t dim, uint32_t topk, uint32_t threads, uint32_t *ids, float *dist) {
  if (threads < 1) threads = 1;
  std::vector<std::thread> th;
  for (uint32_t t = 0; t < threads; t++)
    th.emplace_back([&, t] {
      std::vector<std::pair<float, uint32_t>> d(n);
      for (uint32_t i = t; i < nq; i += threads) {
        for (uint32_t j = 0; j < n; j++) {
          const float *a = q + size_t(i) * dim, *b = data + size_t(j) * dim;
          d[j] = {metric == kL2 ? l2_sqr(a, b, dim) : ip_sqr(a, b, dim), j};
        }
        uint32_t k = std::min(topk, n);
        std::partial_sort(d.begin(), d.begin() + k, d.end());
        for (uint32_t j = 0; j < topk; j++) {
          ids[size_t(i) * topk + j] = j < k ? d[j].second : 0;
          if (dist) dist[size_t(i) * topk + j] = j < k ? d[j].first : 0.f;
        }
      }
    });
  for (auto &x : th) x.join();
  return 0;
}
// include/utils/evaluate.hpp:64-81
float orc_calc_recall(const uint32_t *res, const uint32_t *gt, uint32_t nq, uint32_t gt_dim, uint32_t topk) {
  uint32_t cnt = 0;
  for (uint32_t i = 0; i < nq; i++)
    for (uint32_t j = 0; j < topk; j++)
      for (uint32_t k = 0; k < gt_dim; k++)
        if (res[i * topk + j] == gt[i * gt_dim + k]) {
          cnt++;
          break;
        }
  return static_cast<float>(cnt) / static_cast<float>(nq * topk);
}

// LinearPool replay (same contract as ref_linear_pool_replay)
int orc_linear_pool_replay(uint32_t n, int capacity, const uint32_t *op_ids, const float *op_dist,
                           uint32_t nops, uint32_t *popped, uint32_t *npopped, uint32_t *final_ids,
                           float *final_dist, uint32_t *final_size) {
  Pool pool(n, capacity);
  uint32_t np = 0;
  for (uint32_t i = 0; i < nops; i++) {
    if (op_ids[i] == kEmpty) {
      popped[np++] = pool.pop();  // like the reference's tests: pop() itself does not check has_next()
    } else {
      pool.insert(op_ids[i], op_dist[i]);
    }
  }
  *npopped = np;
  *final_size = static_cast<uint32_t>(pool.size);
  for (size_t i = 0; i < pool.size; i++) {
    final_ids[i] = pool.id(i);
    final_dist[i] = pool.data[i].dist;
  }
  return 0;
}

// HNSW level sequence (hnswlib.hpp:182-186 with seed 100, M=16): first `count` levels
int orc_hnsw_levels(uint32_t count, uint32_t max_edge, uint32_t *out) {
  std::default_random_engine gen;
  gen.seed(100);
  double mult = 1 / std::log(1.0 * static_cast<double>(max_edge));
  for (uint32_t i = 0; i < count; i++) {
    std::uniform_real_distribution<double> distribution(0.0, 1.0);
    double r = -std::log(distribution(gen)) * mult;
    out[i] = static_cast<uint32_t>(static_cast<size_t>(r));
  }
  return 0;
}

}  // extern "C"
