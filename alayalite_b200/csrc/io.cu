// On-disk formats, byte-compatible with the reference so that CPU- and GPU-built indices
// interchange (paths relative to /root/reference):
//   Graph::save / load              include/index/graph/graph.hpp:165-238
//   OverlayGraph::save / load       include/index/graph/overlay_graph.hpp:151-194
//   SequentialStorage::save / load  include/storage/sequential_storage.hpp:110-142
//   RawSpace::save / load           include/space/raw_space.hpp:219-250
//   SQ8Space / SQ4Space save / load include/space/sq8_space.hpp:213-251, sq4_space.hpp:214-254
//   SQ8Quantizer::save / load       include/space/quant/sq8.hpp:160-177
// All values little-endian, written as raw in-memory values (size_t = 8 bytes, IDType = uint32).
// The device keeps rows in exactly the storage layout of the reference (64-byte row stride, zero or
// 0xFF fill), so row blocks are streamed between HBM and the file without repacking.
#include <cstdio>
#include <cstring>
#include <memory>

#include "index.cuh"
#include "kernels.cuh"

namespace ab {

namespace {

struct File {
  FILE *f = nullptr;
  std::string path;
  File(const char *p, const char *mode) : path(p) {
    f = std::fopen(p, mode);
    if (!f) throw Error(AB_ERR_IO, "Cannot open file " + path);
  }
  ~File() {
    if (f) std::fclose(f);
  }
  void write(const void *p, size_t n) {
    if (n && std::fwrite(p, 1, n, f) != n) throw Error(AB_ERR_IO, "short write to " + path);
  }
  void read(void *p, size_t n) {
    if (n && std::fread(p, 1, n, f) != n) throw Error(AB_ERR_IO, "unexpected end of file " + path);
  }
  template <typename T>
  void put(T v) { write(&v, sizeof(T)); }
  template <typename T>
  T get() {
    T v;
    read(&v, sizeof(T));
    return v;
  }
  bool at_eof() {
    int c = std::fgetc(f);
    if (c == EOF) return true;
    std::ungetc(c, f);
    return false;
  }
};

constexpr size_t kIoChunk = 64u << 20;

void write_device_block(File &f, const void *d_ptr, size_t bytes, cudaStream_t s) {
  std::vector<uint8_t> buf(std::min(bytes, kIoChunk));
  for (size_t o = 0; o < bytes; o += buf.size()) {
    const size_t m = std::min(buf.size(), bytes - o);
    AB_CUDA(cudaMemcpyAsync(buf.data(), static_cast<const uint8_t *>(d_ptr) + o, m, cudaMemcpyDeviceToHost, s));
    AB_CUDA(cudaStreamSynchronize(s));
    f.write(buf.data(), m);
  }
}
void read_device_block(File &f, void *d_ptr, size_t bytes, cudaStream_t s) {
  std::vector<uint8_t> buf(std::min(bytes, kIoChunk));
  for (size_t o = 0; o < bytes; o += buf.size()) {
    const size_t m = std::min(buf.size(), bytes - o);
    f.read(buf.data(), m);
    AB_CUDA(cudaMemcpyAsync(static_cast<uint8_t *>(d_ptr) + o, buf.data(), m, cudaMemcpyHostToDevice, s));
    AB_CUDA(cudaStreamSynchronize(s));
  }
}

std::vector<uint8_t> pack_bits(const std::vector<uint8_t> &valid, size_t capacity) {
  std::vector<uint8_t> bm((capacity + 7) / 8, 0);
  for (size_t i = 0; i < capacity && i < valid.size(); i++)
    if (valid[i]) bm[i / 8] |= static_cast<uint8_t>(1u << (i % 8));
  return bm;
}

// SequentialStorage::save (sequential_storage.hpp:110-119)
void write_storage(File &f, size_t item_size, size_t capacity, size_t pos, const void *d_rows,
                   const std::vector<uint8_t> &valid, cudaStream_t s) {
  const size_t aligned = round_up(item_size, kRowAlign);
  f.put<uint64_t>(item_size);
  f.put<uint64_t>(aligned);
  f.put<uint64_t>(capacity);
  f.put<uint64_t>(pos);
  f.put<uint64_t>(kRowAlign);
  write_device_block(f, d_rows, aligned * capacity, s);
  const auto bm = pack_bits(valid, capacity);
  f.write(bm.data(), bm.size());
}

struct StorageHeader {
  uint64_t item_size, aligned, capacity, pos, alignment;
};
StorageHeader read_storage_header(File &f) {
  StorageHeader h;
  h.item_size = f.get<uint64_t>();
  h.aligned = f.get<uint64_t>();
  h.capacity = f.get<uint64_t>();
  h.pos = f.get<uint64_t>();
  h.alignment = f.get<uint64_t>();
  AB_CHECK(h.aligned == round_up(h.item_size, kRowAlign) && h.alignment == kRowAlign, AB_ERR_IO,
           "unsupported storage alignment in " + f.path);
  return h;
}
std::vector<uint8_t> read_bitmap(File &f, size_t capacity) {
  std::vector<uint8_t> bm((capacity + 7) / 8), valid(capacity, 0);
  f.read(bm.data(), bm.size());
  for (size_t i = 0; i < capacity; i++) valid[i] = (bm[i / 8] >> (i % 8)) & 1;
  return valid;
}

void write_space_header(File &f, int metric, uint32_t data_size, uint32_t dim, uint32_t item_cnt, uint32_t delete_cnt,
                        uint32_t capacity) {
  f.put<int32_t>(metric);
  f.put<uint32_t>(data_size);
  f.put<uint32_t>(dim);
  f.put<uint32_t>(item_cnt);
  f.put<uint32_t>(delete_cnt);
  f.put<uint32_t>(capacity);
}

void upload_valid_bitmap(ab_index *ix) {
  const size_t words = (ix->capacity + 31) / 32;
  std::vector<uint32_t> w(words, 0);
  for (size_t i = 0; i < ix->capacity; i++)
    if (ix->h_valid_raw[i]) w[i / 32] |= 1u << (i % 32);
  ix->valid.reserve(words);
  AB_CUDA(cudaMemcpyAsync(ix->valid.p, w.data(), words * 4, cudaMemcpyHostToDevice, ix->stream));
  AB_CUDA(cudaStreamSynchronize(ix->stream));
}

}  // namespace

void index_sync_valid_bitmap(ab_index *ix) { upload_valid_bitmap(ix); }

void index_save(ab_index *ix, const char *index_path, const char *data_path, const char *quant_path) {
  AB_CHECK(ix->has_graph, AB_ERR_STATE, "index is not fitted or loaded");
  cudaStream_t s = ix->stream;
  const uint32_t cap = static_cast<uint32_t>(ix->capacity);
  {
    // Graph::save: nep, eps, max_nodes, max_nbrs, storage, [overlay]
    File f(index_path, "wb");
    if (ix->has_overlay) {
      f.put<int32_t>(0);  // HNSWBuilder never fills eps_ (hnsw_builder.hpp:98-194)
    } else {
      f.put<int32_t>(1);
      f.put<uint32_t>(ix->ep);
    }
    f.put<uint32_t>(cap);
    f.put<uint32_t>(32);
    write_storage(f, 32 * sizeof(uint32_t), cap, ix->n, ix->adj.p, ix->h_valid_graph, s);
    if (ix->has_overlay) {
      std::vector<uint32_t> lists(std::max<uint64_t>(ix->ov_rows, 1) * 32);
      if (ix->ov_rows)
        AB_CUDA(cudaMemcpy(lists.data(), ix->ov_lists.p, ix->ov_rows * 32 * sizeof(uint32_t), cudaMemcpyDeviceToHost));
      std::vector<uint32_t> off(cap);
      AB_CUDA(cudaMemcpy(off.data(), ix->ov_off.p, size_t(cap) * 4, cudaMemcpyDeviceToHost));
      f.put<uint32_t>(cap);  // OverlayGraph(node_num = graph->max_nodes_, max_nbrs = graph->max_nbrs_)
      f.put<uint32_t>(32);
      f.put<uint32_t>(ix->ep);
      for (uint32_t i = 0; i < cap; i++) {
        const int32_t cnt = static_cast<int32_t>(ix->h_levels[i] * 32);
        f.put<int32_t>(cnt);
        if (cnt > 0) f.write(lists.data() + size_t(off[i]) * 32, size_t(cnt) * 4);
      }
    }
  }
  if (data_path != nullptr && data_path[0] != 0) {
    File f(data_path, "wb");
    write_space_header(f, ix->params.metric, ix->dim * 4, ix->dim, static_cast<uint32_t>(ix->n_raw), ix->delete_cnt_raw, cap);
    write_storage(f, size_t(ix->dim) * 4, cap, ix->n_raw, ix->raw.p, ix->h_valid_raw, s);
  }
  if (quant_path != nullptr && quant_path[0] != 0) {
    AB_CHECK(ix->params.quantization_type != AB_QUANT_NONE, AB_ERR_STATE, "quant_path given for an index without SQ space");
    File f(quant_path, "wb");
    write_space_header(f, ix->params.metric, ix->row_bytes, ix->dim, static_cast<uint32_t>(ix->n), ix->delete_cnt_space, cap);
    write_storage(f, ix->row_bytes, cap, ix->n, ix->codes.p, ix->h_valid_space, s);
    std::vector<float> mm(size_t(ix->dim) * 2);
    AB_CUDA(cudaMemcpy(mm.data(), ix->qmin.p, ix->dim * 4, cudaMemcpyDeviceToHost));
    AB_CUDA(cudaMemcpy(mm.data() + ix->dim, ix->qmax.p, ix->dim * 4, cudaMemcpyDeviceToHost));
    f.put<uint32_t>(ix->dim);
    f.write(mm.data(), mm.size() * 4);
  }
}

void index_load(ab_index *ix, const char *index_path, const char *data_path, const char *quant_path) {
  cudaStream_t s = ix->stream;
  AB_CHECK(data_path != nullptr && data_path[0] != 0, AB_ERR_INVALID, "load needs the raw data file (index.hpp:165-166)");
  const int quant = ix->params.quantization_type;
  AB_CHECK(quant == AB_QUANT_NONE || (quant_path != nullptr && quant_path[0] != 0), AB_ERR_INVALID,
           "load needs the quantised data file");
  // a load that fails part-way must not leave a searchable half-overwritten index behind
  ix->has_graph = false;

  // ---- raw space header first: it fixes dim and capacity ----------------------------------------
  File fd(data_path, "rb");
  const int32_t metric = fd.get<int32_t>();
  const uint32_t data_size = fd.get<uint32_t>();
  const uint32_t dim = fd.get<uint32_t>();
  const uint32_t item_cnt = fd.get<uint32_t>();
  ix->delete_cnt_raw = fd.get<uint32_t>();
  const uint32_t capacity = fd.get<uint32_t>();
  AB_CHECK(data_size == dim * 4, AB_ERR_UNSUPPORTED, "only float32 raw spaces are supported");
  AB_CHECK(metric >= AB_METRIC_L2 && metric <= AB_METRIC_COS, AB_ERR_IO, "bad metric in " + fd.path);
  ix->params.metric = metric;
  ix->capacity = capacity;
  ix->params.capacity = capacity;
  index_configure_layout(ix, dim);
  index_alloc_spaces(ix);
  AB_CUDA(cudaStreamSynchronize(s));
  StorageHeader hr = read_storage_header(fd);
  AB_CHECK(hr.item_size == size_t(dim) * 4 && hr.capacity == capacity, AB_ERR_IO, "inconsistent raw storage header");
  read_device_block(fd, ix->raw.p, hr.aligned * hr.capacity, s);
  ix->h_valid_raw = read_bitmap(fd, capacity);
  ix->n_raw = item_cnt;
  ix->n = item_cnt;

  // ---- quantised space -----------------------------------------------------------------------
  if (quant != AB_QUANT_NONE) {
    File fq(quant_path, "rb");
    const int32_t qmetric = fq.get<int32_t>();
    const uint32_t qsize = fq.get<uint32_t>();
    const uint32_t qdim = fq.get<uint32_t>();
    const uint32_t qcnt = fq.get<uint32_t>();
    ix->delete_cnt_space = fq.get<uint32_t>();
    const uint32_t qcap = fq.get<uint32_t>();
    AB_CHECK(qdim == dim && qsize == ix->row_bytes && qcap == capacity, AB_ERR_IO, "quantised space does not match the raw space");
    ix->params.metric = qmetric;  // the search space's own metric drives the traversal
    StorageHeader hq = read_storage_header(fq);
    AB_CHECK(hq.item_size == ix->row_bytes && hq.capacity == capacity, AB_ERR_IO, "inconsistent code storage header");
    read_device_block(fq, ix->codes.p, hq.aligned * hq.capacity, s);
    ix->h_valid_space = read_bitmap(fq, capacity);
    AB_CHECK(fq.get<uint32_t>() == dim, AB_ERR_IO, "quantizer dimension mismatch");
    std::vector<float> mm(size_t(dim) * 2);
    fq.read(mm.data(), mm.size() * 4);
    AB_CUDA(cudaMemcpy(ix->qmin.p, mm.data(), dim * 4, cudaMemcpyHostToDevice));
    AB_CUDA(cudaMemcpy(ix->qmax.p, mm.data() + dim, dim * 4, cudaMemcpyHostToDevice));
    quant_tables(quant, ix->qmin.p, ix->qmax.p, dim, ix->scale.p, ix->vmin.p, ix->table_len, s);
    ix->n = qcnt;
  } else {
    ix->h_valid_space = ix->h_valid_raw;
    ix->delete_cnt_space = ix->delete_cnt_raw;
  }

  // ---- graph -----------------------------------------------------------------------------------
  File fg(index_path, "rb");
  const int32_t nep = fg.get<int32_t>();
  std::vector<uint32_t> eps(std::max(nep, 0));
  fg.read(eps.data(), eps.size() * 4);
  const uint32_t max_nodes = fg.get<uint32_t>();
  const uint32_t max_nbrs = fg.get<uint32_t>();
  AB_CHECK(max_nbrs == 32, AB_ERR_UNSUPPORTED, "only max_nbrs == 32 graphs are supported");
  AB_CHECK(max_nodes == capacity, AB_ERR_IO, "graph capacity does not match the space capacity");
  StorageHeader hg = read_storage_header(fg);
  AB_CHECK(hg.item_size == 128 && hg.capacity == capacity, AB_ERR_IO, "inconsistent graph storage header");
  read_device_block(fg, ix->adj.p, hg.aligned * hg.capacity, s);
  ix->h_valid_graph = read_bitmap(fg, capacity);
  AB_CHECK(hg.pos >= ix->n, AB_ERR_IO, "graph has fewer rows than the search space");
  ix->has_overlay = !fg.at_eof();  // graph.hpp:233-236
  if (ix->has_overlay) {
    const uint32_t node_num = fg.get<uint32_t>();
    const uint32_t ov_nbrs = fg.get<uint32_t>();
    const uint32_t ep = fg.get<uint32_t>();
    AB_CHECK(ov_nbrs == 32 && node_num <= capacity, AB_ERR_IO, "unsupported overlay header");
    std::vector<uint32_t> off(capacity, kEmptyId), lists;
    uint64_t rows = 0;
    for (uint32_t i = 0; i < node_num; i++) {
      const int32_t cnt = fg.get<int32_t>();
      AB_CHECK(cnt >= 0 && cnt % 32 == 0, AB_ERR_IO, "bad overlay list length");
      ix->h_levels[i] = static_cast<uint32_t>(cnt / 32);
      if (cnt > 0) {
        off[i] = static_cast<uint32_t>(rows);
        lists.resize((rows + cnt / 32) * 32);
        fg.read(lists.data() + rows * 32, size_t(cnt) * 4);
        rows += cnt / 32;
      }
    }
    ix->ov_rows = rows;
    ix->ov_lists.reserve(std::max<uint64_t>(rows, 1) * 32);
    if (rows) AB_CUDA(cudaMemcpy(ix->ov_lists.p, lists.data(), rows * 32 * 4, cudaMemcpyHostToDevice));
    AB_CUDA(cudaMemcpy(ix->ov_off.p, off.data(), size_t(capacity) * 4, cudaMemcpyHostToDevice));
    AB_CHECK(ep < ix->n, AB_ERR_IO, "overlay entry point out of range in " + fg.path);
    ix->ep = ep;
    ix->ep_level = ix->h_levels[ep];
  } else {
    AB_CHECK(nep > 0, AB_ERR_IO, "graph has neither overlay nor entry points");
    AB_CHECK(eps[0] < ix->n, AB_ERR_IO, "graph entry point out of range in " + fg.path);
    ix->ep = eps[0];
    ix->ep_level = 0;
    ix->ov_rows = 0;
  }
  // tombstones of the raw space (raw_space.hpp:297-300)
  ix->has_removed = false;
  for (uint64_t i = 0; i < ix->n_raw; i++)
    if (!ix->h_valid_raw[i]) ix->has_removed = true;
  if (ix->has_removed) upload_valid_bitmap(ix);
  ix->has_graph = true;
  AB_CUDA(cudaStreamSynchronize(s));
}

}  // namespace ab
