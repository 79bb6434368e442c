// pybind_module.cpp -- `pydegensac_b200.pydegensac`: the pybind11 surface of the reference
// (bindings.cpp:470-506: findHomography_, findFundamentalMatrix_, same argument names, order and
// defaults) over the C ABI of libdegensac_b200.so.  Additive keyword: seed (the reference seeds from
// time(NULL) and has no seed argument).  The GIL is released while the GPU works.
#include <pybind11/numpy.h>
#include <pybind11/pybind11.h>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/degensac_b200.h"

namespace py = pybind11;

static void raise_for(int rc) {
  if (rc == DGB200_OK) return;
  const std::string msg = dgb200_last_error();
  if (rc == DGB200_E_ARG || rc == DGB200_E_METRIC) throw std::invalid_argument(msg);  // -> ValueError, as bindings.cpp:32-47
  if (rc == DGB200_E_UNSUPPORTED) throw std::invalid_argument(msg);
  throw std::runtime_error("degensac_b200: " + msg);
}

static void check_shapes(const py::buffer_info& a, const py::buffer_info& b, size_t min_n) {
  if (a.ndim != 2 || b.ndim != 2) throw std::invalid_argument("x1y1 should be an array with dims [n,2], [n,6]");
  const size_t n = a.shape[0], dim = a.shape[1], n2 = b.shape[0], dim2 = b.shape[1];
  if ((dim != 2) && (dim != 6)) throw std::invalid_argument("x1y1 should be an array with dims [n,2], [n,6], n>=4");
  if (n < min_n) throw std::invalid_argument("x1y1 should be an array with dims [n,2], n>=" + std::to_string(min_n));
  if ((dim2 != 2) && (dim2 != 6)) throw std::invalid_argument("x2y2 should be an array with dims [n,2] or [n, 6], n>=4");
  if (n2 != n) throw std::invalid_argument("x1y1 and x2y2 should be the same size");
  if (dim2 != dim) throw std::invalid_argument("x1y1 and x2y2 should have the same number of columns");
}

static py::tuple findHomography_(py::array_t<double, py::array::c_style | py::array::forcecast> x1y1,
                                 py::array_t<double, py::array::c_style | py::array::forcecast> x2y2, double px_th,
                                 double conf, int max_iters, int error_type, bool sym_check_enable, double laf_coef,
                                 uint64_t seed) {
  py::buffer_info b1 = x1y1.request(), b2 = x2y2.request();
  check_shapes(b1, b2, 4);
  const int n = (int)b1.shape[0], dim = (int)b1.shape[1];
  py::array_t<double> H_out({3, 3});
  py::array_t<bool> mask_out(n);
  std::vector<uint8_t> mask(n);
  double* model = (double*)H_out.request().ptr;
  int rc;
  {
    py::gil_scoped_release nogil;
    rc = dgb200_find_homography((const double*)b1.ptr, (const double*)b2.ptr, n, dim, px_th, conf, max_iters, error_type,
                                sym_check_enable ? 1 : 0, laf_coef, seed, model, mask.data(), nullptr);
  }
  raise_for(rc);
  bool* m = (bool*)mask_out.request().ptr;
  for (int i = 0; i < n; ++i) m[i] = mask[i] != 0;
  return py::make_tuple(H_out, mask_out);
}

static py::tuple findFundamentalMatrix_(py::array_t<double, py::array::c_style | py::array::forcecast> x1y1,
                                        py::array_t<double, py::array::c_style | py::array::forcecast> x2y2,
                                        double px_th, double conf, int max_iters, int error_type, bool sym_check_enable,
                                        double laf_coef, bool enable_degeneracy_check, uint64_t seed) {
  py::buffer_info b1 = x1y1.request(), b2 = x2y2.request();
  check_shapes(b1, b2, 8);
  const int n = (int)b1.shape[0], dim = (int)b1.shape[1];
  py::array_t<double> F_out({3, 3});
  py::array_t<bool> mask_out(n);
  std::vector<uint8_t> mask(n);
  double* model = (double*)F_out.request().ptr;
  int rc;
  {
    py::gil_scoped_release nogil;
    rc = dgb200_find_fundamental((const double*)b1.ptr, (const double*)b2.ptr, n, dim, px_th, conf, max_iters,
                                 error_type, sym_check_enable ? 1 : 0, laf_coef, enable_degeneracy_check ? 1 : 0, seed,
                                 model, mask.data(), nullptr);
  }
  raise_for(rc);
  bool* m = (bool*)mask_out.request().ptr;
  for (int i = 0; i < n; ++i) m[i] = mask[i] != 0;
  return py::make_tuple(F_out, mask_out);
}

PYBIND11_MODULE(pydegensac, m) {
  m.doc() = "B200-native LO-RANSAC / DEGENSAC (drop-in for pydegensac.pydegensac)";
  m.def("findHomography_", &findHomography_, "LO-RANSAC homography (raw core output: inv(H.T) is applied by utils)",
        py::arg("x1y1"), py::arg("x2y2"), py::arg("px_th") = 1.0, py::arg("conf") = 0.999, py::arg("max_iters") = 10000,
        py::arg("error_type") = 0, py::arg("sym_check_enable") = 1, py::arg("laf_coef") = 0, py::arg("seed") = 0);
  m.def("findFundamentalMatrix_", &findFundamentalMatrix_, "LO-RANSAC / DEGENSAC fundamental matrix", py::arg("x1y1"),
        py::arg("x2y2"), py::arg("px_th") = 0.5, py::arg("conf") = 0.9999, py::arg("max_iters") = 200000,
        py::arg("error_type") = 0, py::arg("sym_check_enable") = 1, py::arg("laf_coef") = 0,
        py::arg("enable_degeneracy_check") = 1, py::arg("seed") = 0);
}
