// PyTorch extension shim over the C ABI (include/eigenpde_b200.h): TORCH_LIBRARY ops that unpack tensors into the plain
// pointers the library takes and launch on at::cuda::getCurrentCUDAStream().  No arithmetic lives here -- the sm_100a kernels
// are in libeigenpde_b200.so; this file is what makes the step a first-class torch operator (dispatcher, stream semantics,
// usable from torch.autograd.Function / C++ frontends) as BASELINE.json's north_star describes.
#include <ATen/ATen.h>
#include <c10/cuda/CUDAGuard.h>
#include <c10/cuda/CUDAStream.h>
#include <torch/library.h>

#include <string>
#include <vector>

#include "../../include/eigenpde_b200.h"

namespace {

epde_desc make_desc(const std::vector<int64_t>& arch, int64_t k, int64_t act_id, int64_t d_pad, int64_t flags,
                    const c10::optional<at::Tensor>& metric) {
  TORCH_CHECK(arch.size() >= 2 && arch.size() <= EPDE_MAX_LAYERS + 1, "arch = [d, n_1, ..., 1]");
  epde_desc D{};
  D.d = (int32_t)arch[0]; D.d_pad = (int32_t)d_pad; D.k = (int32_t)k; D.n_layers = (int32_t)arch.size() - 1;
  for (size_t i = 0; i < arch.size(); i++) D.widths[i] = (int32_t)arch[i];
  D.act_id = (int32_t)act_id; D.flags = (uint32_t)flags;
  D.metric = metric.has_value() ? metric->data_ptr<float>() : nullptr;
  return D;
}
void check(int rc, const char* what) { TORCH_CHECK(rc == 0, what, " failed: ", rc, " (", epde_error_string(rc), ")"); }
void* stream_of(const at::Tensor& t) { return (void*)c10::cuda::getCurrentCUDAStream(t.get_device()).stream(); }

// epde_step: loss + gradient of eigen_solver.update_step (src/pytorch_eigen_solver.py:212-233); fills out (fp64) and grad (fp32)
void step(const at::Tensor& X, const at::Tensor& w, const c10::optional<at::Tensor>& idx, int64_t B, const at::Tensor& params,
          const at::Tensor& diag, double beta, double alpha, std::vector<double> eig_w, std::vector<int64_t> arch, int64_t k,
          int64_t act_id, int64_t flags, const c10::optional<at::Tensor>& metric, at::Tensor ws, at::Tensor out, at::Tensor grad) {
  TORCH_CHECK(X.is_cuda() && X.scalar_type() == at::kFloat && X.is_contiguous(), "X: contiguous fp32 CUDA [K, d_pad]");
  TORCH_CHECK(params.scalar_type() == at::kFloat && grad.scalar_type() == at::kFloat && out.scalar_type() == at::kDouble);
  TORCH_CHECK((int64_t)eig_w.size() >= k, "eig_w shorter than k");
  c10::cuda::CUDAGuard guard(X.device());
  const epde_desc D = make_desc(arch, k, act_id, X.size(1), flags, metric);
  const int64_t* ip = idx.has_value() ? idx->data_ptr<int64_t>() : nullptr;
  check(epde_step(&D, X.data_ptr<float>(), w.data_ptr<float>(), ip, B, params.data_ptr<float>(), diag.data_ptr<float>(), beta,
                  alpha, eig_w.data(), ws.data_ptr(), (size_t)ws.numel(), out.data_ptr<double>(), grad.data_ptr<float>(),
                  stream_of(X)), "epde_step");
}

void step_partial(const at::Tensor& X, const at::Tensor& w, const c10::optional<at::Tensor>& idx, int64_t B,
                  const at::Tensor& params, const at::Tensor& diag, std::vector<int64_t> arch, int64_t k, int64_t act_id,
                  int64_t flags, const c10::optional<at::Tensor>& metric, at::Tensor ws, at::Tensor sums) {
  c10::cuda::CUDAGuard guard(X.device());
  const epde_desc D = make_desc(arch, k, act_id, X.size(1), flags, metric);
  const int64_t* ip = idx.has_value() ? idx->data_ptr<int64_t>() : nullptr;
  check(epde_step_partial(&D, X.data_ptr<float>(), w.data_ptr<float>(), ip, B, params.data_ptr<float>(), diag.data_ptr<float>(),
                          ws.data_ptr(), (size_t)ws.numel(), sums.data_ptr<double>(), stream_of(X)), "epde_step_partial");
}

void step_finish(const at::Tensor& X, const at::Tensor& w, const c10::optional<at::Tensor>& idx, int64_t B,
                 const at::Tensor& params, const at::Tensor& diag, double beta, double alpha, std::vector<double> eig_w,
                 std::vector<int64_t> arch, int64_t k, int64_t act_id, int64_t flags, const c10::optional<at::Tensor>& metric,
                 const at::Tensor& sums, at::Tensor ws, at::Tensor out, at::Tensor grad) {
  c10::cuda::CUDAGuard guard(X.device());
  const epde_desc D = make_desc(arch, k, act_id, X.size(1), flags, metric);
  const int64_t* ip = idx.has_value() ? idx->data_ptr<int64_t>() : nullptr;
  check(epde_step_finish(&D, X.data_ptr<float>(), w.data_ptr<float>(), ip, B, params.data_ptr<float>(), diag.data_ptr<float>(),
                         beta, alpha, eig_w.data(), sums.data_ptr<double>(), ws.data_ptr(), (size_t)ws.numel(),
                         out.data_ptr<double>(), grad.data_ptr<float>(), stream_of(X)), "epde_step_finish");
}

int64_t workspace_bytes(std::vector<int64_t> arch, int64_t k, int64_t act_id, int64_t d_pad, int64_t flags, int64_t B) {
  const epde_desc D = make_desc(arch, k, act_id, d_pad, flags, c10::nullopt);
  size_t need = 0;
  check(epde_workspace_bytes(&D, B, &need), "epde_workspace_bytes");
  return (int64_t)need;
}

}  // namespace

TORCH_LIBRARY(epde, m) {
  m.def("step(Tensor X, Tensor w, Tensor? idx, int B, Tensor params, Tensor diag, float beta, float alpha, float[] eig_w, "
        "int[] arch, int k, int act_id, int flags, Tensor? metric, Tensor(a!) ws, Tensor(b!) out, Tensor(c!) grad) -> ()", &step);
  m.def("step_partial(Tensor X, Tensor w, Tensor? idx, int B, Tensor params, Tensor diag, int[] arch, int k, int act_id, "
        "int flags, Tensor? metric, Tensor(a!) ws, Tensor(b!) sums) -> ()", &step_partial);
  m.def("step_finish(Tensor X, Tensor w, Tensor? idx, int B, Tensor params, Tensor diag, float beta, float alpha, "
        "float[] eig_w, int[] arch, int k, int act_id, int flags, Tensor? metric, Tensor sums, Tensor(a!) ws, Tensor(b!) out, "
        "Tensor(c!) grad) -> ()", &step_finish);
  m.def("workspace_bytes(int[] arch, int k, int act_id, int d_pad, int flags, int B) -> int", &workspace_bytes);
}
