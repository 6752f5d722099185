// dag_loss_fn_shim.cpp -- pybind module `dag_loss_fn_b200`: the reference's native boundary (its module `dag_loss_fn`) on top
// of libdagb200.so.  (A distinct module name: pybind11 caches extension modules by name, so two modules called
// `dag_loss_fn` cannot coexist in one process -- the differential tests load the reference's next to this one.)
//
// The reference binds four torch::Tensor functions in DASpeech/custom_ops/dag_loss.cpp:19-29 and its Python layer calls
// them through get_dag_kernel() (custom_ops/dag_loss.py:37-64, 105, 118, 169, 182, 227, 272).  This file exports the
// same four names with the same argument lists and return values, implemented as thin calls into the C ABI of
// include/dagb200.h (plain device pointers + the CURRENT torch stream).  A maintainer who keeps the reference's
// dag_loss.py unchanged makes get_dag_kernel() return this module instead of JIT-compiling the reference sources
// (option B).  Built in-tree by daspeech_b200/csrc/build_shim.py; no kernels live here.
#include <torch/extension.h>
#include <c10/cuda/CUDAGuard.h>
#include <c10/cuda/CUDAStream.h>

#include <map>
#include <mutex>
#include <tuple>
#include <utility>

#include "../../include/dagb200.h"

namespace {

int dtype_code(const torch::Tensor &t, const char *who) {
  switch (t.scalar_type()) {
    case at::kFloat: return DAGB200_F32;
    case at::kDouble: return DAGB200_F64;
    case at::kHalf: return DAGB200_F16;
    case at::kBFloat16: return DAGB200_BF16;
    default: TORCH_CHECK(false, "\"", who, "\" not implemented for '", t.scalar_type(), "'");
  }
  return -1;
}

void check_rc(int rc, const char *what) { TORCH_CHECK(rc == 0, what, " failed (code ", rc, "): ", dagb200_last_error()); }

// grow-only scratch per (device, stream): the blocked kernels need ~400 MB at the C2 shape and the contents only live
// for one call; calls on a stream are ordered.
void *workspace(size_t nbytes, const torch::Tensor &like, void *stream) {
  if (nbytes == 0) return nullptr;
  static std::mutex mu;
  static std::map<std::pair<int, void *>, torch::Tensor> cache;
  std::lock_guard<std::mutex> lock(mu);
  auto key = std::make_pair((int)like.get_device(), stream);
  auto it = cache.find(key);
  if (it == cache.end() || (size_t)it->second.numel() < nbytes) {
    cache.erase(key);
    it = cache.emplace(key, torch::empty({(int64_t)nbytes}, like.options().dtype(torch::kUInt8))).first;
  }
  return it->second.data_ptr();
}

// the reference's argument checks (dag_loss.cu:317-332, dag_best_alignment.cu:212-227)
void check_lattice(const torch::Tensor &match_all, const torch::Tensor &links, const torch::Tensor &output_length,
                   const torch::Tensor &target_length) {
  TORCH_CHECK(match_all.is_cuda() && links.is_cuda() && output_length.is_cuda() && target_length.is_cuda(),
              "all inputs must be CUDA tensors");
  TORCH_CHECK(match_all.dim() == 3, "match_all dim != 3");
  TORCH_CHECK(links.dim() == 3, "links dim != 3");
  TORCH_CHECK(output_length.dim() == 1, "output_length dim != 3");
  TORCH_CHECK(target_length.dim() == 1, "target_length dim != 3");
  const auto bsz = match_all.size(0);
  TORCH_CHECK(links.size(0) == bsz && output_length.size(0) == bsz && target_length.size(0) == bsz, "batch size not match");
  TORCH_CHECK(links.size(1) == match_all.size(2), "prelen not match");
  TORCH_CHECK(output_length.scalar_type() == at::kLong && target_length.scalar_type() == at::kLong, "length should be long");
  TORCH_CHECK(match_all.scalar_type() == at::kFloat || match_all.scalar_type() == at::kDouble,
              "\"dag_loss\" not implemented for '", match_all.scalar_type(), "'");
  TORCH_CHECK(links.scalar_type() == match_all.scalar_type(), "match_all and links must have the same dtype");
}

}  // namespace

// The four entry points live in their own namespace with internal linkage: the reference's extension exports global
// functions with the same names and signatures (dag_loss.cpp:19-22), and both modules may be loaded in one process.
namespace dagb200_shim {

// dag_loss.cpp:19 / dag_loss.cu:313-375
static std::tuple<torch::Tensor, torch::Tensor> dag_loss(const torch::Tensor &match_all_, const torch::Tensor &links_,
                                                  const torch::Tensor &output_length_, const torch::Tensor &target_length_,
                                                  bool require_gradient, int config) {
  check_lattice(match_all_, links_, output_length_, target_length_);
  const c10::cuda::CUDAGuard guard(match_all_.device());
  auto match_all = match_all_.contiguous(), links = links_.contiguous();
  auto olen = output_length_.contiguous(), tlen = target_length_.contiguous();
  const int B = (int)match_all.size(0), M = (int)match_all.size(1), L = (int)match_all.size(2), T = (int)links.size(2);
  auto alpha = torch::empty_like(match_all), beta = torch::empty_like(match_all);
  void *stream = c10::cuda::getCurrentCUDAStream().stream();
  const size_t nbytes = match_all.scalar_type() == at::kFloat ? dagb200_dag_loss_workspace_bytes(B, M, L, T) : 0;
  check_rc(dagb200_dag_loss(match_all.data_ptr(), links.data_ptr(), olen.data_ptr<int64_t>(), tlen.data_ptr<int64_t>(),
                            alpha.data_ptr(), beta.data_ptr(), dtype_code(match_all, "dag_loss"), B, M, L, T,
                            require_gradient ? 1 : 0, config, workspace(nbytes, match_all, stream), nbytes,
                            /*status=*/nullptr, stream),
           "dag_loss");
  return std::make_tuple(alpha, beta);
}

// dag_loss.cpp:20 / dag_loss.cu:518-571
static std::tuple<torch::Tensor, torch::Tensor> dag_loss_backward(const torch::Tensor &grad_output_, const torch::Tensor &alpha_,
                                                           const torch::Tensor &beta_, const torch::Tensor &match_all_,
                                                           const torch::Tensor &links_, const torch::Tensor &output_length_,
                                                           const torch::Tensor &target_length_, int config1, int config2) {
  const c10::cuda::CUDAGuard guard(match_all_.device());
  auto match_all = match_all_.contiguous(), links = links_.contiguous();
  auto alpha = alpha_.contiguous(), beta = beta_.contiguous();
  auto olen = output_length_.contiguous(), tlen = target_length_.contiguous();
  auto go = grad_output_.to(match_all.scalar_type()).contiguous();
  const int B = (int)match_all.size(0), M = (int)match_all.size(1), L = (int)match_all.size(2), T = (int)links.size(2);
  auto grad_match_all = torch::empty_like(match_all), grad_links = torch::empty_like(links);
  void *stream = c10::cuda::getCurrentCUDAStream().stream();
  const size_t nbytes = match_all.scalar_type() == at::kFloat ? dagb200_dag_loss_backward_workspace_bytes(B, M, L, T) : 0;
  check_rc(dagb200_dag_loss_backward_ws(go.data_ptr(), alpha.data_ptr(), beta.data_ptr(), match_all.data_ptr(),
                                        links.data_ptr(), olen.data_ptr<int64_t>(), tlen.data_ptr<int64_t>(),
                                        grad_match_all.data_ptr(), grad_links.data_ptr(),
                                        dtype_code(match_all, "dag_loss_backward"), B, M, L, T, config1, config2,
                                        workspace(nbytes, match_all, stream), nbytes, stream),
           "dag_loss_backward");
  return std::make_tuple(grad_match_all, grad_links);
}

// dag_loss.cpp:21 / dag_best_alignment.cu:209-253: returns (max-plus lattice, path int32 [B, L])
static std::tuple<torch::Tensor, torch::Tensor> dag_best_alignment(const torch::Tensor &match_all_, const torch::Tensor &links_,
                                                            const torch::Tensor &output_length_,
                                                            const torch::Tensor &target_length_, int config) {
  check_lattice(match_all_, links_, output_length_, target_length_);
  const c10::cuda::CUDAGuard guard(match_all_.device());
  auto match_all = match_all_.contiguous(), links = links_.contiguous();
  auto olen = output_length_.contiguous(), tlen = target_length_.contiguous();
  const int B = (int)match_all.size(0), M = (int)match_all.size(1), L = (int)match_all.size(2), T = (int)links.size(2);
  auto alpha = torch::empty_like(match_all);
  auto path = torch::empty({B, L}, match_all.options().dtype(torch::kInt32));
  void *stream = c10::cuda::getCurrentCUDAStream().stream();
  const size_t nbytes = dagb200_best_alignment_workspace_bytes(B, M, L, T);
  check_rc(dagb200_dag_best_alignment(match_all.data_ptr(), links.data_ptr(), olen.data_ptr<int64_t>(),
                                      tlen.data_ptr<int64_t>(), alpha.data_ptr(), path.data_ptr<int32_t>(),
                                      dtype_code(match_all, "dag_best_alignment"), B, M, L, T, config,
                                      workspace(nbytes > 0 ? nbytes : 1, match_all, stream), nbytes, /*status=*/nullptr, stream),
           "dag_best_alignment");
  return std::make_tuple(alpha, path);
}

// dag_loss.cpp:22 / logsoftmax_gather.cu:313-377: word_ins_out is overwritten with probabilities iff require_gradient
static torch::Tensor logsoftmax_gather(torch::Tensor word_ins_out, const torch::Tensor &select_idx, bool require_gradient) {
  TORCH_CHECK(word_ins_out.is_cuda() && select_idx.is_cuda(), "inputs must be CUDA tensors");
  TORCH_CHECK(word_ins_out.dim() == 3, "word_ins_out dim != 3");
  TORCH_CHECK(select_idx.dim() == 3, "select_idx dim != 3");
  const int B = (int)word_ins_out.size(0), L = (int)word_ins_out.size(1), V = (int)word_ins_out.size(2);
  const int S = (int)select_idx.size(2);
  TORCH_CHECK(select_idx.size(0) == B, "batch size not match");
  TORCH_CHECK(select_idx.size(1) == L, "prelen size not match");
  TORCH_CHECK(select_idx.scalar_type() == at::kLong, "select_idx should be long");
  TORCH_CHECK(word_ins_out.is_contiguous(), "word_ins_out is not contiguous");
  const c10::cuda::CUDAGuard guard(word_ins_out.device());
  const auto out_dtype = word_ins_out.scalar_type() == at::kDouble ? torch::kDouble : torch::kFloat;
  // [B, S, L]-contiguous buffer returned as a [B, L, S] view: the criterion's transpose(1, 2) + .contiguous() are free
  auto result = torch::empty({B, S, L}, word_ins_out.options().dtype(out_dtype)).transpose(1, 2);
  check_rc(dagb200_logsoftmax_gather(word_ins_out.data_ptr(), dtype_code(word_ins_out, "logsoftmax_gather"),
                                     select_idx.data_ptr<int64_t>(), select_idx.stride(0), select_idx.stride(1),
                                     select_idx.stride(2), result.data_ptr(), result.stride(0), result.stride(1),
                                     result.stride(2), B, L, V, S, require_gradient ? 1 : 0,
                                     c10::cuda::getCurrentCUDAStream().stream()),
           "logsoftmax_gather");
  return result;
}

}  // namespace dagb200_shim

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
  using namespace dagb200_shim;
  m.def("dag_loss", &dagb200_shim::dag_loss, "alpha / beta lattices of the DAG log-marginal (libdagb200)");
  m.def("dag_loss_backward", &dagb200_shim::dag_loss_backward, "emission and transition gradients (libdagb200)");
  m.def("dag_best_alignment", &dagb200_shim::dag_best_alignment, "Viterbi lattice and alignment path (libdagb200)");
  m.def("logsoftmax_gather", &dagb200_shim::logsoftmax_gather, "fused vocabulary log-softmax + target gather (libdagb200)");
}
