// pgeof_ext.cpp -- Python binding of the B200-native pgeof hot path.
//
// Mirrors the surface of the reference's nanobind module (src/pgeof_ext.cpp:12-178): the
// same seven function names, keyword names, defaults, strict dtypes ("noconvert": a wrong
// dtype is a TypeError, never a silent cast) and the EFeatureID enum with exported values.
// nanobind is not available in this image, so the TU is written against
// pybind11; it only parses arguments and forwards raw pointers to the C ABI of
// include/pgeof_b200.h -- swapping the binding toolkit does not touch the library.
//
// Inputs may be numpy arrays (host flavour of the ABI; results are numpy arrays backed by
// the library's pinned pool) or CUDA tensors exposing __cuda_array_interface__ / __dlpack__
// (device flavour on torch's current stream; results are torch tensors on the same device).
#include <pybind11/numpy.h>
#include <pybind11/pybind11.h>
#include <pybind11/stl.h>

#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "pgeof_b200.h"

namespace py = pybind11;
using namespace py::literals;

namespace {

enum EFeatureID {   // include/pca.hpp:47-63
    Linearity = 0, Planarity, Scattering, VerticalityPGEOF, Normal_x, Normal_y, Normal_z, Length, Surface, Volume, Curvature,
    K_optimal, Verticality, Eigentropy
};

int g_eig_order = -1;   // -1: follow the PGEOF_EIG_ORDER environment variable

int eig_order()
{
    if (g_eig_order >= 0) return g_eig_order;
    const char* e = std::getenv("PGEOF_EIG_ORDER");
    if (e && (std::strcmp(e, "documented") == 0 || std::strcmp(e, "descending") == 0)) return PGEOF_EIG_DOCUMENTED;
    return PGEOF_EIG_LITERAL;
}

void check(int status)
{
    if (status == PGEOF_OK) return;
    const std::string msg = pgeof_last_error();
    switch (status) {
        case PGEOF_EINVAL: throw py::value_error(msg);          // std::invalid_argument in the reference
        case PGEOF_EINDEX: throw py::index_error(msg);
        case PGEOF_ENOMEM: PyErr_SetString(PyExc_MemoryError, msg.c_str()); throw py::error_already_set();
        default: throw std::runtime_error(msg);
    }
}

// ----------------------------------------------------------------------------------
// argument views
// ----------------------------------------------------------------------------------
struct ArrayView {
    const void* ptr = nullptr;
    std::vector<size_t> shape;
    bool on_device = false;
    py::object keep;     // keeps a (possibly compacted) source alive for the call
    py::object torch_device;   // device of a torch tensor input
    size_t size() const { size_t n = 1; for (size_t s : shape) n *= s; return n; }
};

bool is_numpy(const py::object& o) { return py::isinstance<py::array>(o); }

py::module_ torch_module() { return py::module_::import("torch"); }

bool is_torch_tensor(const py::object& o)
{
    return py::hasattr(o, "data_ptr") && py::hasattr(o, "is_cuda") && py::hasattr(o, "dtype");
}

// typestr of __cuda_array_interface__ / numpy for the dtype a parameter requires
struct DType { const char* name; char kind; int itemsize; };
constexpr DType kF32{"float32", 'f', 4}, kF64{"float64", 'f', 8}, kU32{"uint32", 'u', 4}, kU64{"uint64", 'u', 8}, kI64{"int64", 'i', 8};

[[noreturn]] void type_error(const char* arg, const DType& want, const std::string& got)
{
    throw py::type_error(std::string("argument '") + arg + "' must be a " + want.name +
                         " array (no implicit conversion, as in pgeof); got " + got);
}

ArrayView view_numpy(py::array a, const char* arg, const DType& want, int ndim)
{
    const py::dtype dt = a.dtype();
    if (dt.kind() != want.kind || dt.itemsize() != want.itemsize || !a.dtype().attr("isnative").cast<bool>())
        type_error(arg, want, py::str(dt).cast<std::string>());
    if (a.ndim() != ndim) throw py::type_error(std::string("argument '") + arg + "' must have " + std::to_string(ndim) + " dimension(s)");
    if (ndim == 2 && a.shape(1) != 3) throw py::type_error(std::string("argument '") + arg + "' must have shape (n, 3)");
    if (!(a.flags() & py::array::c_style)) {
        // Eigen::Ref<const RowMajor (n,3)> accepts any row stride with unit inner stride
        // (pca.hpp:15-18); flat index arrays are walked as dense (pgeof.hpp:85-86).
        const bool rows_ok = ndim == 2 && a.strides(1) == (py::ssize_t)want.itemsize && a.strides(0) >= 3 * (py::ssize_t)want.itemsize;
        if (!rows_ok) throw py::type_error(std::string("argument '") + arg + "' must be C-contiguous");
        a = py::array::ensure(a, py::array::c_style);
    }
    ArrayView v;
    v.ptr = a.data();
    for (int i = 0; i < a.ndim(); ++i) v.shape.push_back((size_t)a.shape(i));
    v.keep = a;
    return v;
}

ArrayView view_torch(py::object t, const char* arg, const DType& want, int ndim)
{
    const std::string dt = py::str(t.attr("dtype")).cast<std::string>();   // "torch.float32"
    if (dt != std::string("torch.") + want.name) type_error(arg, want, dt);
    if (t.attr("dim")().cast<int>() != ndim) throw py::type_error(std::string("argument '") + arg + "' must have " + std::to_string(ndim) + " dimension(s)");
    if (!t.attr("is_contiguous")().cast<bool>()) {
        if (ndim != 2) throw py::type_error(std::string("argument '") + arg + "' must be contiguous");
        t = t.attr("contiguous")();
    }
    ArrayView v;
    for (auto s : t.attr("shape")) v.shape.push_back(s.cast<size_t>());
    if (ndim == 2 && v.shape[1] != 3) throw py::type_error(std::string("argument '") + arg + "' must have shape (n, 3)");
    if (!t.attr("is_cuda").cast<bool>()) {   // CPU tensor: share memory with numpy
        py::array a = t.attr("detach")().attr("numpy")();
        return view_numpy(a, arg, want, ndim);
    }
    v.ptr = reinterpret_cast<const void*>(t.attr("data_ptr")().cast<uintptr_t>());
    v.on_device = true;
    v.keep = t;
    v.torch_device = t.attr("device");
    return v;
}

ArrayView view_any(const py::object& o, const char* arg, const DType& want, int ndim)
{
    if (is_numpy(o)) return view_numpy(py::reinterpret_borrow<py::array>(o), arg, want, ndim);
    if (is_torch_tensor(o)) return view_torch(o, arg, want, ndim);
    if (py::hasattr(o, "__dlpack__")) return view_torch(torch_module().attr("from_dlpack")(o), arg, want, ndim);
    if (py::hasattr(o, "__cuda_array_interface__")) return view_torch(torch_module().attr("as_tensor")(o, "device"_a = "cuda"), arg, want, ndim);
    throw py::type_error(std::string("argument '") + arg + "' must be a numpy array or a CUDA tensor (" + want.name + ")");
}

void same_space(std::initializer_list<const ArrayView*> views)
{
    const ArrayView* first = *views.begin();
    for (const ArrayView* v : views) {
        if (v->on_device != first->on_device) throw py::type_error("all array arguments must live in the same memory space (all numpy or all CUDA)");
        if (v->on_device && !v->torch_device.equal(first->torch_device)) throw py::type_error("all CUDA tensors must be on the same device");
    }
}

// ----------------------------------------------------------------------------------
// outputs
// ----------------------------------------------------------------------------------
struct Output {
    py::object obj;
    void* ptr = nullptr;
};

Output make_numpy(const std::vector<size_t>& shape, const char* dtype, size_t itemsize)
{
    size_t n = 1;
    std::vector<py::ssize_t> shp, strides(shape.size());
    for (size_t s : shape) { n *= s; shp.push_back((py::ssize_t)s); }
    py::ssize_t st = (py::ssize_t)itemsize;
    for (size_t i = shape.size(); i-- > 0;) { strides[i] = st; st *= (py::ssize_t)shape[i]; }
    Output o;
    void* pinned = n ? pgeof_host_alloc(n * itemsize) : nullptr;
    if (pinned) {   // pinned result buffer from the library pool, returned to it when numpy drops the array
        py::capsule owner(pinned, [](void* p) { pgeof_host_free(p); });
        o.obj = py::array(py::dtype(dtype), shp, strides, pinned, owner);
        o.ptr = pinned;
    } else {
        py::array a(py::dtype(dtype), shp, strides);
        o.obj = a;
        o.ptr = a.mutable_data();
    }
    return o;
}

Output make_torch(const std::vector<size_t>& shape, const char* dtype, const py::object& device)
{
    py::module_ torch = torch_module();
    py::tuple shp(shape.size());
    for (size_t i = 0; i < shape.size(); ++i) shp[i] = shape[i];
    Output o;
    o.obj = torch.attr("empty")(shp, "dtype"_a = torch.attr(dtype), "device"_a = device);
    o.ptr = reinterpret_cast<void*>(o.obj.attr("data_ptr")().cast<uintptr_t>());
    return o;
}

Output make_output(const ArrayView& like, const std::vector<size_t>& shape, const char* dtype, size_t itemsize)
{
    return like.on_device ? make_torch(shape, dtype, like.torch_device) : make_numpy(shape, dtype, itemsize);
}

// torch's current stream on the tensor's device; the guard makes that device current
struct TorchStream {
    py::object guard;
    void* stream = nullptr;
    explicit TorchStream(const ArrayView& v)
    {
        if (!v.on_device) return;
        py::module_ cuda = torch_module().attr("cuda");
        guard = cuda.attr("device")(v.torch_device);
        guard.attr("__enter__")();
        stream = reinterpret_cast<void*>(cuda.attr("current_stream")(v.torch_device).attr("cuda_stream").cast<uintptr_t>());
    }
    ~TorchStream() { if (guard) guard.attr("__exit__")(py::none(), py::none(), py::none()); }
};

template <typename F>
void run_nogil(F&& f)
{
    int status;
    {
        py::gil_scoped_release release;
        status = f();
    }
    check(status);
}

void say_done(bool verbose)
{
    if (verbose) py::print("100% done");   // the reference prints a racy progress counter (pgeof.hpp:32-48)
}

// ----------------------------------------------------------------------------------
// the seven entry points (src/pgeof_ext.cpp:34-177)
// ----------------------------------------------------------------------------------
py::tuple knn_search(const py::object& data, const py::object& query, uint32_t knn)
{
    ArrayView d = view_any(data, "data", kF32, 2), q = view_any(query, "query", kF32, 2);
    same_space({&d, &q});
    if (knn > d.shape[0]) throw py::value_error("knn size is greater than the data point cloud size");   // nn_search.hpp:37
    Output idx = make_output(q, {q.shape[0], knn}, "uint32", 4), d2 = make_output(q, {q.shape[0], knn}, "float32", 4);
    TorchStream ts(d);
    const float* qp = static_cast<const float*>(q.ptr);
    if (data.is(query)) qp = static_cast<const float*>(d.ptr);
    run_nogil([&] {
        return d.on_device ? pgeof_knn_search_dev((const float*)d.ptr, d.shape[0], qp, q.shape[0], knn, (uint32_t*)idx.ptr, (float*)d2.ptr, ts.stream)
                           : pgeof_knn_search((const float*)d.ptr, d.shape[0], qp, q.shape[0], knn, (uint32_t*)idx.ptr, (float*)d2.ptr);
    });
    return py::make_tuple(idx.obj, d2.obj);
}

py::tuple radius_search(const py::object& data, const py::object& query, float search_radius, uint32_t max_knn)
{
    ArrayView d = view_any(data, "data", kF32, 2), q = view_any(query, "query", kF32, 2);
    same_space({&d, &q});
    if (max_knn > d.shape[0]) throw py::value_error("max knn size is greater than the data point cloud size");   // nn_search.hpp:92-95
    Output idx = make_output(q, {q.shape[0], max_knn}, "int32", 4), d2 = make_output(q, {q.shape[0], max_knn}, "float32", 4);
    TorchStream ts(d);
    const float* qp = static_cast<const float*>(q.ptr);
    if (data.is(query)) qp = static_cast<const float*>(d.ptr);
    run_nogil([&] {
        return d.on_device ? pgeof_radius_search_dev((const float*)d.ptr, d.shape[0], qp, q.shape[0], search_radius, max_knn, (int32_t*)idx.ptr, (float*)d2.ptr, ts.stream)
                           : pgeof_radius_search((const float*)d.ptr, d.shape[0], qp, q.shape[0], search_radius, max_knn, (int32_t*)idx.ptr, (float*)d2.ptr);
    });
    return py::make_tuple(idx.obj, d2.obj);
}

struct Csr {
    ArrayView xyz, nn, ptr;
    size_t n_rows;
    bool p64;     // nn_ptr holds 64-bit offsets (extension; the reference binds uint32 only, pgeof.hpp:78-79)
};

// 8-byte integer offsets: numpy uint64 / int64, torch uint64 / int64 (torch's usual index dtype)
int wide_offsets(const py::object& o)
{
    std::string name;
    if (is_numpy(o)) {
        const py::dtype dt = py::reinterpret_borrow<py::array>(o).dtype();
        if (dt.itemsize() == 8 && dt.kind() == 'u') return 1;
        if (dt.itemsize() == 8 && dt.kind() == 'i') return 2;
        return 0;
    }
    if (py::hasattr(o, "dtype")) name = py::str(o.attr("dtype")).cast<std::string>();
    if (name == "torch.uint64") return 1;
    if (name == "torch.int64") return 2;
    return 0;
}

Csr view_csr(const py::object& xyz, const py::object& nn, const py::object& nn_ptr)
{
    const int wide = wide_offsets(nn_ptr);
    Csr c{view_any(xyz, "xyz", kF32, 2), view_any(nn, "nn", kU32, 1), view_any(nn_ptr, "nn_ptr", wide == 1 ? kU64 : (wide == 2 ? kI64 : kU32), 1), 0, wide != 0};
    same_space({&c.xyz, &c.nn, &c.ptr});
    if (c.ptr.shape[0] == 0) throw py::value_error("nn_ptr must hold at least one offset");   // reference underflows (pgeof.hpp:83)
    c.n_rows = c.ptr.shape[0] - 1;
    return c;
}

py::object compute_features(const py::object& xyz, const py::object& nn, const py::object& nn_ptr, long k_min, bool verbose)
{
    Csr c = view_csr(xyz, nn, nn_ptr);
    if (k_min < 1) throw py::value_error("k_min should be > 1");   // pgeof.hpp:81
    Output out = make_output(c.xyz, {c.n_rows, 11}, "float32", 4);
    TorchStream ts(c.xyz);
    const int order = eig_order();
    run_nogil([&] {
        if (c.p64)
            return c.xyz.on_device
                       ? pgeof_compute_features_p64_dev((const float*)c.xyz.ptr, c.xyz.shape[0], (const uint32_t*)c.nn.ptr, c.nn.shape[0], (const uint64_t*)c.ptr.ptr, c.n_rows, (uint32_t)k_min, order, (float*)out.ptr, ts.stream)
                       : pgeof_compute_features_p64((const float*)c.xyz.ptr, c.xyz.shape[0], (const uint32_t*)c.nn.ptr, c.nn.shape[0], (const uint64_t*)c.ptr.ptr, c.n_rows, (uint32_t)k_min, order, (float*)out.ptr);
        return c.xyz.on_device
                   ? pgeof_compute_features_dev((const float*)c.xyz.ptr, c.xyz.shape[0], (const uint32_t*)c.nn.ptr, c.nn.shape[0], (const uint32_t*)c.ptr.ptr, c.n_rows, (uint32_t)k_min, order, (float*)out.ptr, ts.stream)
                   : pgeof_compute_features((const float*)c.xyz.ptr, c.xyz.shape[0], (const uint32_t*)c.nn.ptr, c.nn.shape[0], (const uint32_t*)c.ptr.ptr, c.n_rows, (uint32_t)k_min, order, (float*)out.ptr);
    });
    say_done(verbose);
    return out.obj;
}

py::object compute_features_multiscale(const py::object& xyz, const py::object& nn, const py::object& nn_ptr, const py::object& k_scales, bool verbose)
{
    Csr c = view_csr(xyz, nn, nn_ptr);
    std::vector<uint32_t> scales;   // any integer sequence (list or numpy, tests/test_pgeof.py:42,44)
    for (auto s : k_scales) scales.push_back(py::reinterpret_borrow<py::object>(s).attr("__index__")().cast<uint32_t>());
    uint32_t prev = 1;
    for (uint32_t s : scales) { if (s < prev) throw py::value_error("k_scales should be > 1 and sorted in ascending order"); prev = s; }
    Output out = make_output(c.xyz, {c.n_rows, scales.size(), 11}, "float32", 4);
    TorchStream ts(c.xyz);
    const int order = eig_order();
    run_nogil([&] {
        if (c.p64)
            return c.xyz.on_device
                       ? pgeof_compute_features_multiscale_p64_dev((const float*)c.xyz.ptr, c.xyz.shape[0], (const uint32_t*)c.nn.ptr, c.nn.shape[0], (const uint64_t*)c.ptr.ptr, c.n_rows, scales.data(), scales.size(), order, (float*)out.ptr, ts.stream)
                       : pgeof_compute_features_multiscale_p64((const float*)c.xyz.ptr, c.xyz.shape[0], (const uint32_t*)c.nn.ptr, c.nn.shape[0], (const uint64_t*)c.ptr.ptr, c.n_rows, scales.data(), scales.size(), order, (float*)out.ptr);
        return c.xyz.on_device
                   ? pgeof_compute_features_multiscale_dev((const float*)c.xyz.ptr, c.xyz.shape[0], (const uint32_t*)c.nn.ptr, c.nn.shape[0], (const uint32_t*)c.ptr.ptr, c.n_rows, scales.data(), scales.size(), order, (float*)out.ptr, ts.stream)
                   : pgeof_compute_features_multiscale((const float*)c.xyz.ptr, c.xyz.shape[0], (const uint32_t*)c.nn.ptr, c.nn.shape[0], (const uint32_t*)c.ptr.ptr, c.n_rows, scales.data(), scales.size(), order, (float*)out.ptr);
    });
    say_done(verbose);
    return out.obj;
}

py::object compute_features_optimal(const py::object& xyz, const py::object& nn, const py::object& nn_ptr, uint32_t k_min, uint32_t k_step,
                                    uint32_t k_min_search, bool verbose)
{
    Csr c = view_csr(xyz, nn, nn_ptr);
    if (k_min < 1 && k_min_search < 1) throw py::value_error("k_min and k_min_search should be > 1");   // pgeof.hpp:250
    if (k_step < 1) throw py::value_error("k_step should be >= 1");
    Output out = make_output(c.xyz, {c.n_rows, 12}, "float32", 4);
    TorchStream ts(c.xyz);
    const int order = eig_order();
    run_nogil([&] {
        if (c.p64)
            return c.xyz.on_device
                       ? pgeof_compute_features_optimal_p64_dev((const float*)c.xyz.ptr, c.xyz.shape[0], (const uint32_t*)c.nn.ptr, c.nn.shape[0], (const uint64_t*)c.ptr.ptr, c.n_rows, k_min, k_step, k_min_search, order, (float*)out.ptr, ts.stream)
                       : pgeof_compute_features_optimal_p64((const float*)c.xyz.ptr, c.xyz.shape[0], (const uint32_t*)c.nn.ptr, c.nn.shape[0], (const uint64_t*)c.ptr.ptr, c.n_rows, k_min, k_step, k_min_search, order, (float*)out.ptr);
        return c.xyz.on_device
                   ? pgeof_compute_features_optimal_dev((const float*)c.xyz.ptr, c.xyz.shape[0], (const uint32_t*)c.nn.ptr, c.nn.shape[0], (const uint32_t*)c.ptr.ptr, c.n_rows, k_min, k_step, k_min_search, order, (float*)out.ptr, ts.stream)
                   : pgeof_compute_features_optimal((const float*)c.xyz.ptr, c.xyz.shape[0], (const uint32_t*)c.nn.ptr, c.nn.shape[0], (const uint32_t*)c.ptr.ptr, c.n_rows, k_min, k_step, k_min_search, order, (float*)out.ptr);
    });
    say_done(verbose);
    return out.obj;
}

bool is_float64(const py::object& o)
{
    if (is_numpy(o)) { const py::dtype dt = py::reinterpret_borrow<py::array>(o).dtype(); return dt.kind() == 'f' && dt.itemsize() == 8; }
    if (py::hasattr(o, "dtype")) return py::str(o.attr("dtype")).cast<std::string>() == "torch.float64";
    return false;
}

py::object compute_features_selected(const py::object& xyz, double search_radius, uint32_t max_knn, const py::object& selected_features)
{
    std::vector<int32_t> ids;
    for (auto f : selected_features) {
        py::object o = py::reinterpret_borrow<py::object>(f);
        if (py::isinstance<EFeatureID>(o)) ids.push_back((int32_t)o.cast<EFeatureID>());
        else ids.push_back(o.attr("__index__")().cast<int32_t>());
    }
    const int order = eig_order();
    if (is_float64(xyz)) {   // the double overload is registered first in the reference (pgeof_ext.cpp:148)
        ArrayView x = view_any(xyz, "xyz", kF64, 2);
        Output out = make_output(x, {x.shape[0], ids.size()}, "float64", 8);
        TorchStream ts(x);
        run_nogil([&] {
            return x.on_device ? pgeof_compute_features_selected_f64_dev((const double*)x.ptr, x.shape[0], search_radius, max_knn, ids.data(), ids.size(), order, (double*)out.ptr, ts.stream)
                               : pgeof_compute_features_selected_f64((const double*)x.ptr, x.shape[0], search_radius, max_knn, ids.data(), ids.size(), order, (double*)out.ptr);
        });
        return out.obj;
    }
    ArrayView x = view_any(xyz, "xyz", kF32, 2);
    Output out = make_output(x, {x.shape[0], ids.size()}, "float32", 4);
    TorchStream ts(x);
    run_nogil([&] {
        return x.on_device ? pgeof_compute_features_selected_f32_dev((const float*)x.ptr, x.shape[0], (float)search_radius, max_knn, ids.data(), ids.size(), order, (float*)out.ptr, ts.stream)
                           : pgeof_compute_features_selected_f32((const float*)x.ptr, x.shape[0], (float)search_radius, max_knn, ids.data(), ids.size(), order, (float*)out.ptr);
    });
    return out.obj;
}

// ----------------------------------------------------------------------------------
// extensions (clearly separate from the drop-in surface)
// ----------------------------------------------------------------------------------
py::tuple radius_search_csr(const py::object& data, const py::object& query, float search_radius, uint32_t max_knn)
{
    ArrayView d = view_any(data, "data", kF32, 2), q = view_any(query, "query", kF32, 2);
    same_space({&d, &q});
    if (max_knn > d.shape[0]) throw py::value_error("max knn size is greater than the data point cloud size");
    Output ptr = make_output(q, {q.shape[0] + 1}, "uint32", 4);
    TorchStream ts(d);
    const float* qp = data.is(query) ? (const float*)d.ptr : (const float*)q.ptr;
    uint64_t nnz = 0;
    run_nogil([&] {
        return d.on_device ? pgeof_radius_search_csr_dev((const float*)d.ptr, d.shape[0], qp, q.shape[0], search_radius, max_knn, (uint32_t*)ptr.ptr, nullptr, &nnz, ts.stream)
                           : pgeof_radius_search_csr((const float*)d.ptr, d.shape[0], qp, q.shape[0], search_radius, max_knn, (uint32_t*)ptr.ptr, nullptr, &nnz);
    });
    Output nn = make_output(q, {(size_t)nnz}, "uint32", 4);
    if (nnz)
        run_nogil([&] {
            return d.on_device ? pgeof_radius_search_csr_dev((const float*)d.ptr, d.shape[0], qp, q.shape[0], search_radius, max_knn, (uint32_t*)ptr.ptr, (uint32_t*)nn.ptr, &nnz, ts.stream)
                               : pgeof_radius_search_csr((const float*)d.ptr, d.shape[0], qp, q.shape[0], search_radius, max_knn, (uint32_t*)ptr.ptr, (uint32_t*)nn.ptr, &nnz);
        });
    return py::make_tuple(nn.obj, ptr.obj);
}

// kNN emitting CSR: (nn uint32 (n * knn), nn_ptr uint32 -- or int64 / uint64 when the CSR holds more than 2^32-1 neighbours or on request)
py::tuple knn_search_csr(const py::object& data, const py::object& query, uint32_t knn, bool wide_offsets_requested)
{
    ArrayView d = view_any(data, "data", kF32, 2), q = view_any(query, "query", kF32, 2);
    same_space({&d, &q});
    if (knn > d.shape[0]) throw py::value_error("knn size is greater than the data point cloud size");
    const bool wide = wide_offsets_requested || (uint64_t)q.shape[0] * knn > 0xffffffffull;
    Output nn = make_output(q, {q.shape[0] * (size_t)knn}, "uint32", 4);
    // torch indexes with int64 and has little uint64 support: wide offsets are int64 on tensors, uint64 on numpy arrays
    Output ptr = wide ? make_output(q, {q.shape[0] + 1}, q.on_device ? "int64" : "uint64", 8) : make_output(q, {q.shape[0] + 1}, "uint32", 4);
    TorchStream ts(d);
    const float* qp = data.is(query) ? (const float*)d.ptr : (const float*)q.ptr;
    run_nogil([&] {
        return d.on_device ? pgeof_knn_search_csr_dev((const float*)d.ptr, d.shape[0], qp, q.shape[0], knn, (uint32_t*)nn.ptr, ptr.ptr, wide ? 64 : 32, ts.stream)
                           : pgeof_knn_search_csr((const float*)d.ptr, d.shape[0], qp, q.shape[0], knn, (uint32_t*)nn.ptr, ptr.ptr, wide ? 64 : 32);
    });
    return py::make_tuple(nn.obj, ptr.obj);
}

py::object knn_features(const py::object& xyz, uint32_t knn, uint32_t k_min, bool return_neighbors)
{
    ArrayView x = view_any(xyz, "xyz", kF32, 2);
    if (knn > x.shape[0]) throw py::value_error("knn size is greater than the data point cloud size");
    Output feat = make_output(x, {x.shape[0], 11}, "float32", 4);
    Output idx, d2;
    if (return_neighbors) { idx = make_output(x, {x.shape[0], knn}, "uint32", 4); d2 = make_output(x, {x.shape[0], knn}, "float32", 4); }
    TorchStream ts(x);
    const int order = eig_order();
    run_nogil([&] {
        return x.on_device ? pgeof_knn_features_dev((const float*)x.ptr, x.shape[0], knn, k_min, order, (uint32_t*)idx.ptr, (float*)d2.ptr, (float*)feat.ptr, ts.stream)
                           : pgeof_knn_features((const float*)x.ptr, x.shape[0], knn, k_min, order, (uint32_t*)idx.ptr, (float*)d2.ptr, (float*)feat.ptr);
    });
    if (return_neighbors) return py::make_tuple(feat.obj, idx.obj, d2.obj);
    return feat.obj;
}

// rows (int64, input order) and coordinates of slab `rank` of `world` along `axis` of a replicated CUDA cloud
py::tuple slab_select(const py::object& xyz, int rank, int world, int axis)
{
    ArrayView x = view_any(xyz, "xyz", kF32, 2);
    if (!x.on_device) throw py::type_error("slab_select takes a CUDA tensor (host clouds: point_geometric_features_b200.shard.slab_queries)");
    TorchStream ts(x);
    pgeof_slab_plan plan;
    run_nogil([&] { return pgeof_slab_plan_dev((const float*)x.ptr, x.shape[0], rank, world, axis, &plan, ts.stream); });
    Output rows = make_torch({(size_t)plan.count}, "int64", x.torch_device), q = make_torch({(size_t)plan.count, 3}, "float32", x.torch_device);
    run_nogil([&] { return pgeof_slab_fill_dev((const float*)x.ptr, x.shape[0], &plan, (int64_t*)rows.ptr, (float*)q.ptr, ts.stream); });
    return py::make_tuple(rows.obj, q.obj);
}

}  // namespace

PYBIND11_MODULE(pgeof_ext, m)
{
    m.doc() = "Compute, for each point in a 3D point cloud, local geometric features -- B200-native CUDA implementation "
              "of the pgeof hot path (drop-in for pgeof.pgeof_ext)";
    py::enum_<EFeatureID>(m, "EFeatureID")
        .value("Linearity", Linearity).value("Planarity", Planarity).value("Scattering", Scattering)
        .value("VerticalityPGEOF", VerticalityPGEOF).value("Normal_x", Normal_x).value("Normal_y", Normal_y)
        .value("Normal_z", Normal_z).value("Length", Length).value("Surface", Surface).value("Volume", Volume)
        .value("Curvature", Curvature).value("K_optimal", K_optimal).value("Verticality", Verticality)
        .value("Eigentropy", Eigentropy)
        .export_values();

    m.def("compute_features", &compute_features, "xyz"_a.noconvert(), "nn"_a.noconvert(), "nn_ptr"_a.noconvert(), "k_min"_a = 1, "verbose"_a = false,
          "Compute the 11 geometric features of every CSR neighbourhood (nn, nn_ptr) -> float32 (num_points, 11).");
    m.def("compute_features_multiscale", &compute_features_multiscale, "xyz"_a.noconvert(), "nn"_a.noconvert(), "nn_ptr"_a.noconvert(), "k_scales"_a,
          "verbose"_a = false, "Features on the first k neighbours for every k in k_scales -> float32 (num_points, n_scales, 11).");
    m.def("compute_features_optimal", &compute_features_optimal, "xyz"_a.noconvert(), "nn"_a.noconvert(), "nn_ptr"_a.noconvert(), "k_min"_a = 1,
          "k_step"_a = 1, "k_min_search"_a = 1, "verbose"_a = false,
          "Features at the neighbourhood size of lowest eigenentropy (Weinmann 2015) -> float32 (num_points, 12), column 11 = k_optimal.");
    m.def("knn_search", &knn_search, "data"_a.noconvert(), "query"_a.noconvert(), "knn"_a,
          "k nearest data points of every query -> (uint32 indices (n, knn), float32 squared distances (n, knn)).");
    m.def("radius_search", &radius_search, "data"_a.noconvert(), "query"_a.noconvert(), "search_radius"_a, "max_knn"_a,
          "max_knn nearest points within search_radius -> (int32 indices padded with -1, float32 squared distances padded with 0).");
    m.def("compute_features_selected", &compute_features_selected, "xyz"_a.noconvert(), "search_radius"_a, "max_knn"_a, "selected_features"_a,
          "Fused radius search + the selected EFeatureID features, float32 or float64 by the dtype of xyz -> (num_points, n_features).");

    // extensions
    m.def("radius_search_csr", &radius_search_csr, "data"_a.noconvert(), "query"_a.noconvert(), "search_radius"_a, "max_knn"_a,
          "Radius search emitting CSR directly -> (nn, nn_ptr) uint32.");
    m.def("knn_search_csr", &knn_search_csr, "data"_a.noconvert(), "query"_a.noconvert(), "knn"_a, "wide_offsets"_a = false,
          "kNN emitting CSR directly -> (nn uint32, nn_ptr uint32; 64-bit offsets beyond 2^32-1 neighbours or with wide_offsets=True).");
    m.def("knn_features", &knn_features, "xyz"_a.noconvert(), "knn"_a, "k_min"_a = 1, "return_neighbors"_a = false,
          "knn_search(xyz, xyz, knn) + compute_features in one device-resident call.");
    m.def("slab_select", &slab_select, "xyz"_a.noconvert(), "rank"_a, "world"_a, "axis"_a = 2,
          "Rows (int64, input order) and coordinates of spatial query shard `rank` of `world` along `axis` -> (rows, query).");
    m.def("set_eig_order", [](const std::string& s) {
        if (s == "literal") g_eig_order = PGEOF_EIG_LITERAL;
        else if (s == "documented") g_eig_order = PGEOF_EIG_DOCUMENTED;
        else if (s == "env") g_eig_order = -1;
        else throw py::value_error("eig_order must be 'literal', 'documented' or 'env'");
    }, "order"_a, "Eigenvalue slot order: 'literal' (pca.hpp:79-89 as written, increasing), 'documented' (decreasing) or 'env' (PGEOF_EIG_ORDER).");
    m.def("get_eig_order", [] { return std::string(eig_order() == PGEOF_EIG_LITERAL ? "literal" : "documented"); });
    m.def("device_count", [] { return pgeof_device_count(); });
    m.def("launch_count", [] { return pgeof_launch_count(); });
    m.def("reset_launch_count", [] { pgeof_reset_launch_count(); });
    m.def("trim", [] { check(pgeof_trim()); });
    m.def("profile_enable", [](bool on) { pgeof_profile_enable(on ? 1 : 0); }, "on"_a);
    m.def("profile_reset", [] { pgeof_profile_reset(); });
    m.def("profile_read", [](const std::string& name) {
        double ms = 0; uint64_t n = 0;
        check(pgeof_profile_read(name.c_str(), &ms, &n));
        return py::make_tuple(ms, n);
    }, "name"_a, "(total device ms, launches) of one hot kernel since the last profile_reset().");
    m.attr("abi_version") = pgeof_abi_version();
}
