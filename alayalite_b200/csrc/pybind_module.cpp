// alayalite_b200._alayalitepy -- the pybind11 host shell, same module surface as the reference's
// alayalite._alayalitepy (/root/reference/python/src/pybind.cpp:37-147), with every method body
// re-pointed at the C ABI of libalaya_b200.so (include/alaya_b200.h). No arithmetic happens here.
//
// Type dispatch (reference: python/include/dispatch.hpp:25-141): vectors of any supported dtype are
// converted to float32 on the way in -- the reference's kernels convert non-float inputs to float
// element by element anyway (include/simd/distance_l2.ipp:729-742) -- and ids are widened to uint64
// on the way out when id_type is uint64.
#include <pybind11/numpy.h>
#include <pybind11/pybind11.h>
#include <pybind11/stl.h>

#include <cstdint>
#include <cstdlib>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <vector>

#include "alaya_b200.h"

namespace py = pybind11;

namespace {

enum class IndexType { FLAT = 0, HNSW = 1, NSG = 2, FUSION = 3 };
enum class MetricType { L2 = 0, IP = 1, COS = 2, NONE = 3 };
enum class QuantizationType { NONE = 0, SQ8 = 1, SQ4 = 2, RABITQ = 3 };

// python/include/params.hpp:29-52
struct IndexParams {
  IndexType index_type_ = IndexType::HNSW;
  py::dtype data_type_ = py::dtype::of<float>();
  py::dtype id_type_ = py::dtype::of<uint32_t>();
  QuantizationType quantization_type_ = QuantizationType::NONE;
  MetricType metric_ = MetricType::L2;
  uint32_t capacity_ = 100000;
  uint32_t max_nbrs_ = 32;

  IndexParams() = default;
  IndexParams(IndexType it, py::dtype dt, py::dtype idt, QuantizationType qt, MetricType mt, uint32_t cap)
      : index_type_(it), data_type_(std::move(dt)), id_type_(std::move(idt)), quantization_type_(qt), metric_(mt), capacity_(cap) {}
};

void check(int rc) {
  if (rc != AB_OK) throw std::runtime_error(ab_last_error());
}

using FloatArray = py::array_t<float, py::array::c_style | py::array::forcecast>;

class PyIndexInterface {
 public:
  explicit PyIndexInterface(const IndexParams &params) : params_(params) {
    // dispatch.hpp:27-47 / :73-91: supported vector and id dtypes
    const char kind = params.data_type_.kind();
    const auto size = params.data_type_.itemsize();
    const bool ok_dtype = (kind == 'f' && (size == 4 || size == 8)) || ((kind == 'i' || kind == 'u') && (size == 1 || size == 4));
    if (!ok_dtype) throw std::runtime_error("Unsupported data type");
    if (!(params.id_type_.kind() == 'u' && (params.id_type_.itemsize() == 4 || params.id_type_.itemsize() == 8)))
      throw std::runtime_error("Unsupported ID type");
    if (params.quantization_type_ == QuantizationType::RABITQ)
      throw std::runtime_error("Unsupported quantization type: RABITQ is outside the B200 hot path");
    ab_params p{};
    p.index_type = static_cast<int32_t>(params.index_type_);
    p.quantization_type = static_cast<int32_t>(params.quantization_type_);
    p.metric = static_cast<int32_t>(params.metric_);
    p.capacity = params.capacity_;
    p.max_nbrs = params.max_nbrs_;
    p.device = -1;
    // Several GPUs without touching the reference's Python API: ALAYA_B200_DEVICES="0,1,2,3" (CUDA ordinals) and
    // optionally ALAYA_B200_SHARDS=S (default 1: every GPU holds a replica and the query batch is split; S > 1: the
    // rows are split by contiguous id ranges over S GPUs and the lists are merged in the search kernel). The index
    // then lives in an ab_group (include/alaya_b200.h, "Multi-GPU"); fit / batch_search work as usual, the
    // single-vector and persistence calls are refused.
    if (const char *devs = std::getenv("ALAYA_B200_DEVICES")) {
      std::vector<int32_t> ids;
      std::stringstream ss(devs);
      for (std::string tok; std::getline(ss, tok, ',');)
        if (!tok.empty()) ids.push_back(std::stoi(tok));
      const char *sh = std::getenv("ALAYA_B200_SHARDS");
      const uint32_t shards = sh ? static_cast<uint32_t>(std::max(1, std::atoi(sh))) : 1u;
      if (ids.size() > 1) {
        if (ids.size() % shards != 0) throw std::runtime_error("ALAYA_B200_DEVICES must hold a multiple of ALAYA_B200_SHARDS devices");
        ab_group *g = nullptr;
        check(ab_group_create(&p, ids.data(), shards, static_cast<uint32_t>(ids.size() / shards), &g));
        group_.reset(g, ab_group_destroy);
        return;
      }
      if (ids.size() == 1) p.device = ids[0];
    }
    ab_index *h = nullptr;
    check(ab_index_create(&p, &h));
    handle_.reset(h, ab_index_destroy);
  }

  void single_gpu_only(const char *what) const {
    if (group_) throw std::runtime_error(std::string(what) + " is not available on a multi-GPU index (ALAYA_B200_DEVICES)");
  }

  auto to_string() -> std::string { return "PyIndexInterface"; }

  void fit(py::array &vectors, uint32_t ef_construction, uint32_t num_threads) {
    if (vectors.ndim() != 2) throw std::runtime_error("Array must be 2D");  // index.hpp:180-182
    FloatArray v = FloatArray::ensure(vectors);
    if (!v) throw std::runtime_error("Unsupported data type");
    const float *ptr = v.data();
    const auto n = static_cast<uint64_t>(v.shape(0));
    const auto d = static_cast<uint32_t>(v.shape(1));
    {
      py::gil_scoped_release nogil;
      if (group_) check(ab_group_fit(group_.get(), ptr, n, d, ef_construction));
      else check(ab_index_fit(handle_.get(), ptr, n, d, ef_construction, num_threads));
    }
    data_dim_ = d;
  }

  auto search(py::array &query, uint32_t topk, uint32_t ef) -> py::array {
    FloatArray q = FloatArray::ensure(query);
    if (!q) throw std::runtime_error("Unsupported data type");
    py::array_t<uint32_t> ids(static_cast<py::ssize_t>(topk));
    if (group_) check(ab_group_batch_search(group_.get(), q.data(), 1, topk, ef, ids.mutable_data(), nullptr));
    else check(ab_index_search(handle_.get(), q.data(), topk, ef, ids.mutable_data()));
    return cast_ids(ids);
  }

  auto get_data_by_id(uint32_t id) -> py::array {
    single_gpu_only("get_data_by_id");
    const uint32_t d = ab_index_get_dim(handle_.get());
    py::array_t<float> out(static_cast<py::ssize_t>(d));
    check(ab_index_get_data_by_id(handle_.get(), id, out.mutable_data()));
    if (params_.data_type_.is(py::dtype::of<float>())) return std::move(out);
    return out.attr("astype")(params_.data_type_).cast<py::array>();
  }

  auto insert(py::array &insert_data, uint32_t ef) -> py::object {
    single_gpu_only("insert");
    FloatArray v = FloatArray::ensure(insert_data);
    if (!v) throw std::runtime_error("Unsupported data type");
    uint32_t id = 0;
    check(ab_index_insert(handle_.get(), v.data(), ef, &id));
    if (id == 0xFFFFFFFFu && params_.id_type_.itemsize() == 8) return py::int_(0xFFFFFFFFFFFFFFFFull);  // (IDType)-1
    return py::int_(id);
  }

  void remove(uint32_t id) {
    single_gpu_only("remove");
    check(ab_index_remove(handle_.get(), id));
  }

  auto batch_search(py::array &queries, uint32_t topk, uint32_t ef, uint32_t num_threads) -> py::array {
    FloatArray q = FloatArray::ensure(queries);
    if (!q || q.ndim() != 2) throw std::runtime_error("Array must be 2D");
    const auto nq = static_cast<uint64_t>(q.shape(0));
    py::array_t<uint32_t> ids(std::vector<py::ssize_t>{static_cast<py::ssize_t>(nq), static_cast<py::ssize_t>(topk)});
    const float *qp = q.data();
    uint32_t *ip = ids.mutable_data();
    {
      py::gil_scoped_release nogil;
      if (group_) check(ab_group_batch_search(group_.get(), qp, nq, topk, ef, ip, nullptr));
      else check(ab_index_batch_search(handle_.get(), qp, nq, topk, ef, num_threads, ip));
    }
    return cast_ids(ids);
  }

  auto batch_search_with_distance(py::array &queries, uint32_t topk, uint32_t ef, uint32_t num_threads) -> py::object {
    FloatArray q = FloatArray::ensure(queries);
    if (!q || q.ndim() != 2) throw std::runtime_error("Array must be 2D");
    const auto nq = static_cast<py::ssize_t>(q.shape(0));
    py::array_t<uint32_t> ids(std::vector<py::ssize_t>{nq, static_cast<py::ssize_t>(topk)});
    py::array_t<float> dist(std::vector<py::ssize_t>{nq, static_cast<py::ssize_t>(topk)});
    int32_t has = 0;
    if (group_) {
      // the merged lists of a group are the re-ranked ones and always carry their exact distances
      check(ab_group_batch_search(group_.get(), q.data(), static_cast<uint64_t>(nq), topk, ef, ids.mutable_data(), dist.mutable_data()));
      has = 1;
    } else {
      check(ab_index_batch_search_with_distance(handle_.get(), q.data(), static_cast<uint64_t>(nq), topk, ef, num_threads,
                                                ids.mutable_data(), dist.mutable_data(), &has));
    }
    // index.hpp:416-418: for SQ spaces dist_pool stays empty -> a (0, topk) array
    if (!has) dist = py::array_t<float>(std::vector<py::ssize_t>{0, static_cast<py::ssize_t>(topk)});
    return py::make_tuple(cast_ids(ids), dist);
  }

  void save(const std::string &index_path, const std::string &data_path, const std::string &quant_path) {
    single_gpu_only("save");
    check(ab_index_save(handle_.get(), index_path.c_str(), data_path.c_str(), quant_path.c_str()));
  }
  void load(const std::string &index_path, const std::string &data_path, const std::string &quant_path) {
    single_gpu_only("load");
    check(ab_index_load(handle_.get(), index_path.c_str(), data_path.c_str(), quant_path.c_str()));
    data_dim_ = ab_index_get_dim(handle_.get());
  }
  auto get_data_dim() -> uint32_t { return data_dim_; }

  // not part of the reference surface: raw handle for the ctypes-level tools (bench, sharding)
  auto handle() -> uintptr_t {
    single_gpu_only("_handle");
    return reinterpret_cast<uintptr_t>(handle_.get());
  }

  IndexParams params_;

 private:
  auto cast_ids(py::array_t<uint32_t> &ids) -> py::array {
    if (params_.id_type_.itemsize() == 8) return ids.attr("astype")(params_.id_type_).cast<py::array>();
    return std::move(ids);
  }
  std::shared_ptr<ab_index> handle_;
  std::shared_ptr<ab_group> group_;  // set instead of handle_ when ALAYA_B200_DEVICES names several GPUs
  uint32_t data_dim_{0};
};

// python/include/client.hpp (unused by the SDK, kept for surface parity)
class Client {
 public:
  auto create_index(const std::string &name, const IndexParams &param) -> std::shared_ptr<PyIndexInterface> {
    if (indices_.count(name)) throw std::runtime_error("Index name already exists");
    auto idx = std::make_shared<PyIndexInterface>(param);
    indices_[name] = idx;
    return idx;
  }
  auto load_index(const std::string &name, const IndexParams &param, const std::string &index_path,
                  const std::string &data_path, const std::string &quant_path) -> std::shared_ptr<PyIndexInterface> {
    if (indices_.count(name)) throw std::runtime_error("Index name already exists");
    auto idx = std::make_shared<PyIndexInterface>(param);
    idx->load(index_path, data_path, quant_path);
    indices_[name] = idx;
    return idx;
  }

 private:
  std::unordered_map<std::string, std::shared_ptr<PyIndexInterface>> indices_;
};

}  // namespace

PYBIND11_MODULE(_alayalitepy, m) {
  m.doc() = "AlayaLite hot path on B200 (libalaya_b200.so)";
  m.attr("__version__") = ab_version();

  py::enum_<IndexType>(m, "IndexType")
      .value("FLAT", IndexType::FLAT)
      .value("HNSW", IndexType::HNSW)
      .value("NSG", IndexType::NSG)
      .value("FUSION", IndexType::FUSION)
      .export_values();
  py::enum_<MetricType>(m, "MetricType")
      .value("L2", MetricType::L2)
      .value("IP", MetricType::IP)
      .value("COS", MetricType::COS)
      .export_values();
  py::enum_<QuantizationType>(m, "QuantizationType")
      .value("NONE", QuantizationType::NONE)
      .value("SQ8", QuantizationType::SQ8)
      .value("SQ4", QuantizationType::SQ4)
      .value("RABITQ", QuantizationType::RABITQ)
      .export_values();

  py::class_<IndexParams>(m, "IndexParams")
      .def(py::init<>())
      .def(py::init<IndexType, py::dtype, py::dtype, QuantizationType, MetricType, uint32_t>(),
           py::arg("index_type_") = IndexType::HNSW, py::arg("data_type_") = py::dtype::of<float>(),
           py::arg("id_type_") = py::dtype::of<uint32_t>(), py::arg("quantization_type_") = QuantizationType::NONE,
           py::arg("metric_") = MetricType::L2, py::arg("capacity_") = 100000u)
      .def_readwrite("index_type_", &IndexParams::index_type_)
      .def_readwrite("data_type_", &IndexParams::data_type_)
      .def_readwrite("id_type_", &IndexParams::id_type_)
      .def_readwrite("quantization_type_", &IndexParams::quantization_type_)
      .def_readwrite("metric_", &IndexParams::metric_)
      .def_readwrite("capacity_", &IndexParams::capacity_);

  py::class_<Client>(m, "Client")
      .def(py::init<>())
      .def("create_index", &Client::create_index, py::arg("name"), py::arg("param"))
      .def("load_index", &Client::load_index, py::arg("name"), py::arg("param"), py::arg("index_path"),
           py::arg("data_path") = std::string(), py::arg("quant_path") = std::string());

  py::class_<PyIndexInterface, std::shared_ptr<PyIndexInterface>>(m, "PyIndexInterface")
      .def(py::init<IndexParams>(), py::arg("params"))
      .def("to_string", &PyIndexInterface::to_string)
      .def("fit", &PyIndexInterface::fit, py::arg("vectors"), py::arg("ef_construction"), py::arg("num_threads"))
      .def("search", &PyIndexInterface::search, py::arg("query"), py::arg("topk"), py::arg("ef"))
      .def("get_data_by_id", &PyIndexInterface::get_data_by_id, py::arg("id"))
      .def("insert", &PyIndexInterface::insert, py::arg("insert_data"), py::arg("ef"))
      .def("remove", &PyIndexInterface::remove, py::arg("id"))
      .def("batch_search", &PyIndexInterface::batch_search, py::arg("queries"), py::arg("topk"), py::arg("ef"),
           py::arg("num_threads"))
      .def("batch_search_with_distance", &PyIndexInterface::batch_search_with_distance, py::arg("queries"),
           py::arg("topk"), py::arg("ef"), py::arg("num_threads"))
      .def("save", &PyIndexInterface::save, py::arg("index_path"), py::arg("data_path"), py::arg("quant_path") = std::string())
      .def("load", &PyIndexInterface::load, py::arg("index_path"), py::arg("data_path"), py::arg("quant_path") = std::string())
      .def("get_data_dim", &PyIndexInterface::get_data_dim)
      .def("_handle", &PyIndexInterface::handle);
}
