// torch_ext.cpp -- thin PyTorch C++ extension over the C ABI (include/pgb200_ert.h): the Python-facing door the
// north star names.  No arithmetic lives here: tensors in, pointers + sizes to libpgb200_ert.so, tensors out, all work
// enqueued on torch's current CUDA stream.  Ops (namespace pgb200):
//   open(pos, node_marker, cells, cell_marker, bounds, bound_marker, sensors, abmn, k?, dim, sr, multilevel, device) -> handle
//   response(handle, model) -> rhoa            model / rhoa: float64 CUDA tensors (inputs already resident in HBM)
//   create_jacobian(handle, model) -> J        zero-copy (rows, cols) view of the column-major HBM buffer
//   jac_mult(handle, x) / jac_tmult(handle, y) host-vector products on the HBM-resident J
//   set_solver(handle, tol, max_iter, check_every), stats(handle) -> float64[15], close(handle)
// Reference boundary mirrored: pygimli/physics/ert/ertModelling.py:213 (response), :238 (createJacobian).
#include <ATen/ATen.h>
#include <torch/library.h>
#include <c10/cuda/CUDAStream.h>
#include <c10/cuda/CUDAGuard.h>

#include "../../include/pgb200_ert.h"

namespace {

inline pgb200_ert *H(int64_t h) { return reinterpret_cast<pgb200_ert *>(static_cast<intptr_t>(h)); }
inline void check(int rc) { TORCH_CHECK(rc == 0, "pgb200: ", pgb200_last_error()); }
inline void on_current_stream(pgb200_ert *h) { check(pgb200_ert_set_stream(h, (void *)c10::cuda::getCurrentCUDAStream().stream())); }

int64_t op_open(const at::Tensor &pos, const at::Tensor &node_marker, const at::Tensor &cells, const at::Tensor &cell_marker,
                const at::Tensor &bounds, const at::Tensor &bound_marker, const at::Tensor &sensors, const at::Tensor &abmn,
                const c10::optional<at::Tensor> &k, int64_t dim, bool sr, bool multilevel, int64_t device) {
    auto P = pos.to(at::kCPU, at::kDouble).contiguous(), S = sensors.to(at::kCPU, at::kDouble).contiguous();
    auto NM = node_marker.to(at::kCPU, at::kInt).contiguous(), CE = cells.to(at::kCPU, at::kInt).contiguous();
    auto CM = cell_marker.to(at::kCPU, at::kInt).contiguous(), B = bounds.to(at::kCPU, at::kInt).contiguous();
    auto BM = bound_marker.to(at::kCPU, at::kInt).contiguous(), AB = abmn.to(at::kCPU, at::kInt).contiguous();
    TORCH_CHECK(P.dim() == 2 && P.size(1) == 3 && CE.dim() == 2 && AB.dim() == 2 && AB.size(1) == 4, "pgb200::open: bad shapes");
    pgb200_mesh_in m{};
    m.dim = (int)dim; m.nloc = (int)CE.size(1); m.n_nodes = (int)P.size(0); m.n_cells = (int)CE.size(0);
    m.n_bounds = (int)BM.numel(); m.nlb = m.n_bounds ? (int)B.size(1) : (int)dim;
    m.pos = P.data_ptr<double>(); m.node_marker = NM.data_ptr<int>(); m.cells = CE.data_ptr<int>(); m.cell_marker = CM.data_ptr<int>();
    m.bounds = B.data_ptr<int>(); m.bound_marker = BM.data_ptr<int>();
    pgb200_scheme_in s{};
    s.n_elec = (int)S.size(0); s.n_data = (int)AB.size(0); s.sensors = S.data_ptr<double>(); s.abmn = AB.data_ptr<int>();
    at::Tensor K;
    if (k.has_value() && k->defined() && k->numel() > 0) { K = k->to(at::kCPU, at::kDouble).contiguous(); s.k_fac = K.data_ptr<double>(); }
    pgb200_ert *h = nullptr;
    check(pgb200_ert_open(&m, &s, sr ? 1 : 0, 0, nullptr, nullptr, multilevel ? 1 : 0, (int)device, &h));
    return (int64_t) reinterpret_cast<intptr_t>(h);
}

at::Tensor op_response(int64_t handle, const at::Tensor &model) {
    TORCH_CHECK(model.is_cuda() && model.scalar_type() == at::kDouble && model.is_contiguous(), "pgb200::response: model must be a contiguous float64 CUDA tensor");
    c10::cuda::CUDAGuard guard(model.device());
    pgb200_ert *h = H(handle);
    on_current_stream(h);
    int rows = 0;
    {   // number of data from the plan
        const pgb200_built_plan *bp = pgb200_ert_plan(h);
        double d = 0.0;
        TORCH_CHECK(bp && pgb200_plan_scalar(bp, "D", &d) == 0, "pgb200::response: handle was not opened through pgb200_ert_open");
        rows = (int)d;
    }
    at::Tensor rhoa = at::empty({rows}, model.options());
    check(pgb200_ert_response_dev(h, model.data_ptr<double>(), (int)model.numel(), rhoa.data_ptr<double>()));
    return rhoa;
}

at::Tensor op_create_jacobian(int64_t handle, const at::Tensor &model) {
    TORCH_CHECK(model.is_cuda() && model.scalar_type() == at::kDouble && model.is_contiguous(), "pgb200::create_jacobian: model must be a contiguous float64 CUDA tensor");
    c10::cuda::CUDAGuard guard(model.device());
    pgb200_ert *h = H(handle);
    on_current_stream(h);
    check(pgb200_ert_create_jacobian_dev(h, model.data_ptr<double>(), (int)model.numel()));
    void *ptr = nullptr; int rows = 0, cols = 0; long long ld = 0;
    check(pgb200_ert_jacobian_info(h, &ptr, &rows, &cols, &ld));
    // column-major [cols][ld] in HBM -> logical (rows, cols) view, no copy; the handle owns the memory
    return at::from_blob(ptr, {rows, cols}, {1, (int64_t)ld}, model.options());
}

at::Tensor op_jac_mult(int64_t handle, const at::Tensor &x) {
    auto X = x.to(at::kCPU, at::kDouble).contiguous();
    void *ptr = nullptr; int rows = 0, cols = 0; long long ld = 0;
    check(pgb200_ert_jacobian_info(H(handle), &ptr, &rows, &cols, &ld));
    TORCH_CHECK(X.numel() == cols, "pgb200::jac_mult: vector length must equal cols");
    at::Tensor y = at::zeros({rows}, X.options());
    check(pgb200_ert_jacobian_mult(H(handle), X.data_ptr<double>(), y.data_ptr<double>()));
    return y;
}
at::Tensor op_jac_tmult(int64_t handle, const at::Tensor &yv) {
    auto Y = yv.to(at::kCPU, at::kDouble).contiguous();
    void *ptr = nullptr; int rows = 0, cols = 0; long long ld = 0;
    check(pgb200_ert_jacobian_info(H(handle), &ptr, &rows, &cols, &ld));
    TORCH_CHECK(Y.numel() == rows, "pgb200::jac_tmult: vector length must equal rows");
    at::Tensor x = at::zeros({cols}, Y.options());
    check(pgb200_ert_jacobian_tmult(H(handle), Y.data_ptr<double>(), x.data_ptr<double>()));
    return x;
}
void op_set_solver(int64_t handle, double tol, int64_t max_iter, int64_t check_every) { check(pgb200_ert_set_solver(H(handle), tol, (int)max_iter, (int)check_every)); }
at::Tensor op_stats(int64_t handle) {
    at::Tensor s = at::zeros({15}, at::TensorOptions().dtype(at::kDouble));
    check(pgb200_ert_stats(H(handle), s.data_ptr<double>(), 15));
    return s;
}
void op_close(int64_t handle) { check(pgb200_ert_destroy(H(handle))); }

} // namespace

TORCH_LIBRARY(pgb200, m) {
    m.def("open(Tensor pos, Tensor node_marker, Tensor cells, Tensor cell_marker, Tensor bounds, Tensor bound_marker, Tensor sensors, Tensor abmn, Tensor? k, int dim, bool sr, bool multilevel, int device) -> int", &op_open);
    m.def("response(int handle, Tensor model) -> Tensor", &op_response);
    m.def("create_jacobian(int handle, Tensor model) -> Tensor", &op_create_jacobian);
    m.def("jac_mult(int handle, Tensor x) -> Tensor", &op_jac_mult);
    m.def("jac_tmult(int handle, Tensor y) -> Tensor", &op_jac_tmult);
    m.def("set_solver(int handle, float tol, int max_iter, int check_every) -> ()", &op_set_solver);
    m.def("stats(int handle) -> Tensor", &op_stats);
    m.def("close(int handle) -> ()", &op_close);
}
