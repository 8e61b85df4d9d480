// Compiled Python binding of the hot entry points of include/mixgrpo_b200.h (pybind11 + libtorch, built in-tree by
// mixgrpo_b200/_build.py into _lib/_torchbind.so).  It is a BINDING, not a second implementation: every function
// validates its tensors, allocates outputs through torch's caching allocator, looks up torch's current CUDA stream and
// calls the same extern "C" entry point the ctypes layer (_cabi.py) calls — mixgrpo_flow_step / _dance_step / _dpm_step,
// mixgrpo_logprob_bwd, mixgrpo_policy_fwd / _bwd — in libmixgrpo_b200.so.  What it removes is interpreter time: one
// eager flow_grpo_step cost 26-36 us of Python + ctypes marshalling for a 7-10 us kernel.
//
// The differentiable transition log-prob (grpo_one_step's operator call, TR:149-168 + loss.backward(), TR:585; TR =
// /root/reference/fastvideo/train_grpo_flux.py) is a torch::autograd::Function here, so forward and backward never
// re-enter the interpreter: forward = fused log-prob kernel on the stored transition, backward = the closed-form
// d log_prob / d model_output kernel (csrc/bwd_kernels.cu).
#include <torch/extension.h>

#include <c10/cuda/CUDAGraphsC10Utils.h>
#include <c10/cuda/CUDAGuard.h>
#include <c10/cuda/CUDAStream.h>

#include <cstring>
#include <string>
#include <tuple>
#include <unordered_map>

#include "mixgrpo_b200.h"

namespace py = pybind11;
using at::Tensor;
using OptT = c10::optional<Tensor>;

namespace {

constexpr int kFlow = 0, kDance = 1, kDpm = 2;

[[noreturn]] void fail_runtime(const std::string& m) { throw std::runtime_error("mixgrpo_b200: " + m); }
[[noreturn]] void fail_value(const std::string& m) { throw py::value_error("mixgrpo_b200: " + m); }
[[noreturn]] void fail_type(const std::string& m) { throw py::type_error("mixgrpo_b200: " + m); }

inline void require_cuda(const Tensor& t, const char* name) {
  if (!t.defined() || !t.is_cuda())
    fail_runtime(std::string("`") + name + "` must be a CUDA tensor — this package has no CPU fallback (got " +
                 (t.defined() ? t.device().str() : std::string("undefined")) + ")");
}

inline int dtype_code(const Tensor& t, const char* name) {
  if (t.scalar_type() == at::kFloat) return MIXGRPO_F32;
  if (t.scalar_type() == at::kBFloat16) return MIXGRPO_BF16;
  fail_type(std::string("`") + name + "` must be float32 or bfloat16, got " + std::string(c10::toString(t.scalar_type())));
}

// dims 1.. form one contiguous block (what ops._rows checks with t[0].is_contiguous())
inline bool inner_contiguous(const Tensor& t) {
  int64_t expect = 1;
  for (int64_t d = t.dim() - 1; d >= 1; --d) {
    if (t.size(d) != 1 && t.stride(d) != expect) return false;
    expect *= t.size(d);
  }
  return true;
}

// (tensor, batch stride in elements) for a (B, ...) tensor; anything not row-contiguous is copied once
inline std::pair<Tensor, int64_t> rows(Tensor t, const char* name) {
  if (t.dim() < 1) fail_value(std::string("`") + name + "` needs a batch dimension");
  if (t.dim() == 1) t = t.unsqueeze(1);
  const int64_t n = t.size(0) ? t.numel() / t.size(0) : 0;
  const bool inner = inner_contiguous(t);
  if (t.size(0) == 1) return {inner ? t : t.contiguous(), n};
  if (inner) return {t, t.stride(0)};
  t = t.contiguous();
  return {t, t.stride(0)};
}

inline Tensor as_f32(const Tensor& t) { return t.scalar_type() == at::kFloat ? t : t.to(at::kFloat); }

void check_rc(int rc, const char* what) {
  if (rc != 0) fail_runtime(std::string(what) + " failed with code " + std::to_string(rc) + ": " + mixgrpo_error_string(rc));
}

// one zero-initialised workspace per (device, stream): the kernels leave the records zeroed, so it is reused by every
// launch on that stream; never cached when allocated inside a CUDA-graph capture (its memory belongs to the graph's pool)
std::unordered_map<uint64_t, Tensor> g_workspaces;
int64_t g_autograd_backward_launches = 0;   // launches issued from TransitionLogProb::backward (the interpreter never sees them)

Tensor workspace(const c10::Device& dev, int64_t B, int64_t n, cudaStream_t st) {
  const int64_t need = mixgrpo_step_workspace_bytes(B, n);
  const uint64_t key = (static_cast<uint64_t>(dev.index()) << 56) ^ reinterpret_cast<uint64_t>(st);
  auto it = g_workspaces.find(key);
  if (it != g_workspaces.end() && it->second.numel() >= need) return it->second;
  Tensor ws = at::zeros({std::max<int64_t>(need, 1 << 16)}, at::TensorOptions().dtype(at::kByte).device(dev));
  if (c10::cuda::currentStreamCaptureStatusMayInitCtx() == c10::cuda::CaptureStatus::None) g_workspaces[key] = ws;
  return ws;
}

inline unsigned make_flags(bool rnd, int vd, int early) {
  return ((rnd && vd == MIXGRPO_BF16) ? MIXGRPO_FLAG_ROUND_LIKE_TORCH : 0u) |
         (early == 2 ? MIXGRPO_FLAG_PDL_EARLY_LOADS : (early == 1 ? MIXGRPO_FLAG_PDL_EARLY_V : 0u));
}

struct StepOut {
  Tensor x_next, x0, logp, mean;   // undefined = not requested
};

// ops.fused_step: one fused sampler step + log-prob launch
StepOut fused_step_impl(int family, Tensor v, Tensor x, const mixgrpo_step_coefs& k, int src, const OptT& noise_o, const OptT& x_next_o,
                        const OptT& m1_o, const OptT& m2_o, int order, bool sde_solver, const OptT& out_x_next, bool want_x0, bool want_mean,
                        bool want_logp, bool rnd, const OptT& out_logp, const OptT& out_x0, int early, bool has_philox, uint64_t ph_seed,
                        uint64_t ph_offset, uint64_t ph_state, const OptT& decode_out, double divisor, double shift, bool from_x0,
                        bool reciprocal, uint64_t defer_ptr = 0, int64_t defer_bytes = 0, const OptT& seed_out = c10::nullopt) {
  require_cuda(v, "model_output");
  require_cuda(x, "latents");
  if (seed_out) {   // trajectory seed: x stays bf16 (the initial latent itself), its fp32 widening is written to seed_out as well
    if (decode_out || x.scalar_type() != at::kBFloat16 || !x.is_contiguous() || seed_out->scalar_type() != at::kFloat ||
        seed_out->sizes() != x.sizes() || !inner_contiguous(*seed_out))
      fail_value("seed_out needs a contiguous bf16 `latents`, an fp32 view of the same shape, and no decode output");
  } else {
    x = as_f32(x);
  }
  const int vd = dtype_code(v, "model_output");
  if (v.sizes() != x.sizes()) fail_value("model_output and latents differ in shape");
  if (!v.is_contiguous()) v = v.contiguous();
  const int64_t B = v.size(0);
  const c10::Device dev = v.device();
  const auto f32 = v.options().dtype(at::kFloat);
  StepOut o;
  if (B == 0) {   // every reference op is a no-op on empty tensors; the per-sample log-prob is an empty vector
    o.x_next = src == MIXGRPO_SRC_GIVEN ? *x_next_o : (out_x_next ? *out_x_next : at::empty(v.sizes(), f32));
    if (want_x0) o.x0 = at::empty(v.sizes(), f32);
    if (want_logp) o.logp = out_logp ? *out_logp : at::empty({0}, f32);
    if (want_mean) o.mean = at::empty(v.sizes(), f32);
    return o;
  }
  const int64_t n = v.numel() / B;
  int64_t x_bs = n;
  if (!seed_out) std::tie(x, x_bs) = rows(x, "latents");
  const void* noise_p = nullptr;
  const float *in_p = nullptr, *m1_p = nullptr, *m2_p = nullptr;
  int64_t in_bs = n;
  Tensor noise, x_next, m1, m2;
  mixgrpo_philox_args pa;
  if (src == MIXGRPO_SRC_NOISE) {
    if (!noise_o) fail_value("rollout step needs explicit `noise`");
    noise = *noise_o;
    require_cuda(noise, "noise");
    const auto want = family == kFlow ? v.scalar_type() : at::kFloat;   // SU:193 vs SU:238 / SU:320
    if (noise.scalar_type() != want || !noise.is_contiguous()) noise = noise.to(want).contiguous();
    if (noise.sizes() != v.sizes()) fail_value("noise shape mismatch");
    noise_p = noise.data_ptr();
  } else if (src == MIXGRPO_SRC_PHILOX) {
    if (!has_philox) fail_value("in-kernel noise needs philox=(seed, offset)");
    pa.seed = ph_seed; pa.offset = ph_offset; pa.device_state = reinterpret_cast<decltype(pa.device_state)>(ph_state);
    noise_p = &pa;                                                      // HOST pointer, read during the call
  } else if (src == MIXGRPO_SRC_GIVEN) {
    if (!x_next_o) fail_value("prev_sample is required");
    x_next = *x_next_o;
    require_cuda(x_next, "prev_sample");
    x_next = as_f32(x_next);
    if (x_next.sizes() != v.sizes()) fail_value("prev_sample shape mismatch");
    std::tie(x_next, in_bs) = rows(x_next, "prev_sample");
    in_p = x_next.data_ptr<float>();
  }
  if (family == kDpm && order >= 2) {
    if (!m1_o) fail_value("dpm order >= 2 needs the previous x0");
    m1 = m1_o->contiguous();
    m1_p = m1.data_ptr<float>();
    if (order == 3) {
      if (!m2_o) fail_value("dpm order 3 needs two previous x0");
      m2 = m2_o->contiguous();
      m2_p = m2.data_ptr<float>();
    }
  }
  float* out_p = nullptr;
  int64_t out_bs = n;
  if (src != MIXGRPO_SRC_GIVEN) {
    if (!out_x_next) {
      o.x_next = at::empty(v.sizes(), f32);
    } else {
      const Tensor& t = *out_x_next;
      if (t.scalar_type() != at::kFloat || t.sizes() != v.sizes() || !inner_contiguous(t)) fail_value("out_x_next must be fp32, same shape, contiguous per sample");
      o.x_next = t;
      if (B > 1) out_bs = t.stride(0);
    }
    out_p = o.x_next.data_ptr<float>();
  } else {
    o.x_next = *x_next_o;
  }
  if (want_x0) {
    if (out_x0) {
      if (out_x0->scalar_type() != at::kFloat || out_x0->sizes() != v.sizes() || !out_x0->is_contiguous())
        fail_value("out_x0 must be a contiguous fp32 tensor of the model output's shape");
      o.x0 = *out_x0;
    } else {
      o.x0 = at::empty(v.sizes(), f32);
    }
  }
  if (want_mean) o.mean = at::empty(v.sizes(), f32);
  if (want_logp) {
    if (out_logp) {
      if (out_logp->scalar_type() != at::kFloat || out_logp->numel() != B || !out_logp->is_contiguous()) fail_value("out_logp must be a contiguous fp32 [B] tensor");
      o.logp = *out_logp;
    } else {
      o.logp = at::empty({B}, f32);
    }
  }
  c10::cuda::OptionalCUDAGuard guard(dev);
  cudaStream_t st = c10::cuda::getCurrentCUDAStream(dev.index()).stream();
  Tensor ws;
  if (want_logp && !defer_ptr) ws = workspace(dev, B, n, st);
  // MIXGRPO_FLAG_DEFER_LOGP: accumulate into the caller's records for this launch; mixgrpo_logp_finalize writes the log-probs
  const unsigned flags = make_flags(rnd, vd, early) | (defer_ptr ? MIXGRPO_FLAG_DEFER_LOGP : 0u);
  mixgrpo_step_ext xe;
  const mixgrpo_step_ext* ext = nullptr;
  if (decode_out) {
    const Tensor& d = *decode_out;
    require_cuda(d, "decode['out']");
    if (d.scalar_type() != at::kFloat || d.dim() != 4 || d.size(0) != B || d.numel() / B != n || !d.is_contiguous())
      fail_value("decode['out'] must be a contiguous fp32 (B, C, H, W) tensor with C*H*W == elements per sample");
    xe.decode_out = d.data_ptr<float>(); xe.C = (int)d.size(1); xe.H = (int)d.size(2); xe.W = (int)d.size(3);
    xe.divisor = (float)divisor; xe.shift = (float)shift; xe.from_x0 = from_x0 ? 1 : 0; xe.reciprocal = reciprocal ? 1 : 0;
    xe.x_is_bf16 = 0; xe.x_f32_out = nullptr; xe.x_f32_out_bs = 0;
    ext = &xe;
  } else if (seed_out) {
    xe.decode_out = nullptr; xe.C = xe.H = xe.W = 0; xe.divisor = 1.f; xe.shift = 0.f; xe.from_x0 = 0; xe.reciprocal = 0;
    xe.x_is_bf16 = 1; xe.x_f32_out = seed_out->data_ptr<float>(); xe.x_f32_out_bs = B > 1 ? seed_out->stride(0) : n;
    ext = &xe;
  }
  float* x0_p = want_x0 ? o.x0.data_ptr<float>() : nullptr;
  float* mean_p = want_mean ? o.mean.data_ptr<float>() : nullptr;
  float* lp_p = want_logp ? o.logp.data_ptr<float>() : nullptr;
  void* ws_p = defer_ptr ? reinterpret_cast<void*>(defer_ptr) : (want_logp ? ws.data_ptr() : nullptr);
  const int64_t ws_n = defer_ptr ? defer_bytes : (want_logp ? ws.numel() : 0);
  int rc;
  if (family == kFlow)
    rc = mixgrpo_flow_step(v.data_ptr(), vd, reinterpret_cast<const float*>(x.data_ptr()), x_bs, noise_p, in_p, in_bs, out_p, out_bs, x0_p, mean_p, lp_p, ws_p, ws_n, B, n, &k,
                           src, flags, st, ext);
  else if (family == kDance)
    rc = mixgrpo_dance_step(v.data_ptr(), vd, reinterpret_cast<const float*>(x.data_ptr()), x_bs, static_cast<const float*>(noise_p), in_p, in_bs, out_p, out_bs, x0_p, mean_p,
                            lp_p, ws_p, ws_n, B, n, &k, src, sde_solver ? 1 : 0, flags, st, ext);
  else if (family == kDpm)
    rc = mixgrpo_dpm_step(v.data_ptr(), vd, reinterpret_cast<const float*>(x.data_ptr()), x_bs, static_cast<const float*>(noise_p), m1_p, m2_p, order, out_p, out_bs, x0_p, mean_p,
                          lp_p, ws_p, ws_n, B, n, &k, src, flags, st, ext);
  else
    fail_value("unknown operator family");
  check_rc(rc, family == kFlow ? "flow_step" : (family == kDance ? "dance_step" : "dpm_step"));
  return o;
}

// ops.logprob_backward: grad of sum_b grad_logp[b]*logp[b] w.r.t. model_output; dtype = model_output.dtype
Tensor logprob_backward_impl(int family, Tensor v, Tensor x, Tensor x_next, const Tensor& grad_logp, const mixgrpo_step_coefs& k, bool rnd,
                             const OptT& out) {
  require_cuda(v, "model_output");
  require_cuda(x, "latents");
  require_cuda(x_next, "prev_sample");
  require_cuda(grad_logp, "grad_log_prob");
  const int vd = dtype_code(v, "model_output");
  if (!v.is_contiguous()) v = v.contiguous();
  if (v.size(0) == 0) return out ? *out : at::empty_like(v);
  const int64_t B = v.size(0), n = v.numel() / B;
  int64_t x_bs, in_bs;
  std::tie(x, x_bs) = rows(as_f32(x), "latents");
  std::tie(x_next, in_bs) = rows(as_f32(x_next), "prev_sample");
  Tensor g = as_f32(grad_logp).contiguous();
  if (g.numel() != B) fail_value("grad_log_prob must have one entry per sample");
  Tensor grad_v = out ? *out : at::empty_like(v);
  if (grad_v.scalar_type() != v.scalar_type() || grad_v.sizes() != v.sizes() || !grad_v.is_contiguous())
    fail_value("`out` must match model_output's dtype/shape and be contiguous");
  c10::cuda::OptionalCUDAGuard guard(v.device());
  cudaStream_t st = c10::cuda::getCurrentCUDAStream(v.device().index()).stream();
  check_rc(mixgrpo_logprob_bwd(family, v.data_ptr(), vd, x.data_ptr<float>(), x_bs, x_next.data_ptr<float>(), in_bs, g.data_ptr<float>(),
                               grad_v.data_ptr(), B, n, &k, make_flags(rnd, vd, 0), st),
           "logprob_bwd");
  return grad_v;
}

inline const mixgrpo_step_coefs& coefs_at(uint64_t addr) {
  if (addr == 0) fail_value("null per-step coefficient block");
  return *reinterpret_cast<const mixgrpo_step_coefs*>(addr);
}

// log p(x_next | x, v) with the closed-form gradient w.r.t. v — sampling_utils._TransitionLogProb without the interpreter
struct TransitionLogProb : public torch::autograd::Function<TransitionLogProb> {
  static torch::autograd::variable_list forward(torch::autograd::AutogradContext* ctx, const Tensor& v, const Tensor& x, const Tensor& x_next,
                                                int64_t coefs_addr, int64_t family, bool rnd, bool want_mean) {
    require_cuda(v, "model_output");                                             // no CPU fallback: fail before anything else is touched
    require_cuda(x, "latents");
    const mixgrpo_step_coefs k = coefs_at(static_cast<uint64_t>(coefs_addr));     // copied: the Python object may die before backward
    StepOut o = fused_step_impl((int)family, v, x, k, MIXGRPO_SRC_GIVEN, c10::nullopt, x_next, c10::nullopt, c10::nullopt, 1, true, c10::nullopt, true,
                                want_mean, true, rnd, c10::nullopt, c10::nullopt, 0, false, 0, 0, 0, c10::nullopt, 1.0, 0.0, false, false);
    ctx->save_for_backward({v, x, x_next});
    ctx->saved_data["k"] = std::string(reinterpret_cast<const char*>(&k), sizeof(k));
    ctx->saved_data["family"] = family;
    ctx->saved_data["rnd"] = rnd;
    Tensor mean = want_mean ? o.mean : at::empty({0}, o.x0.options());
    ctx->mark_non_differentiable({o.x0, mean});
    return {o.logp, o.x0, mean};
  }

  static torch::autograd::variable_list backward(torch::autograd::AutogradContext* ctx, torch::autograd::variable_list grads) {
    const auto saved = ctx->get_saved_variables();
    Tensor grad_v;
    if (ctx->needs_input_grad(0) && grads[0].defined()) {
      mixgrpo_step_coefs k;
      const std::string& raw = ctx->saved_data["k"].toStringRef();
      std::memcpy(&k, raw.data(), sizeof(k));
      grad_v = logprob_backward_impl((int)ctx->saved_data["family"].toInt(), saved[0], saved[1], saved[2], grads[0], k, ctx->saved_data["rnd"].toBool(),
                                     c10::nullopt);
      if (saved[0].size(0) > 0) ++g_autograd_backward_launches;
    }
    return {grad_v, Tensor(), Tensor(), Tensor(), Tensor(), Tensor(), Tensor()};
  }
};

py::object opt(const Tensor& t) { return t.defined() ? py::cast(t) : py::none(); }

struct LossArgsIn {
  Tensor old_lp, adv;
  mixgrpo_loss_args la;
};

LossArgsIn loss_args(const Tensor& old_logp, const Tensor& advantages, const OptT& stats_rows, double clip, double amax, double klc, double denom,
                     int64_t B, const c10::Device& dev, bool accumulate) {
  LossArgsIn r;
  auto vec = [&](const Tensor& t, const char* name) {
    require_cuda(t, name);
    Tensor u = as_f32(t.detach()).contiguous().view({-1});
    if (u.numel() != B) fail_value(std::string("`") + name + "` must have one entry per sample");
    return u;
  };
  r.old_lp = vec(old_logp, "old_log_probs");
  r.adv = vec(advantages, "advantages");
  r.la.old_logp = r.old_lp.data_ptr<float>();
  r.la.advantages = r.adv.data_ptr<float>();
  r.la.stats_rows = nullptr;
  if (stats_rows) {
    const Tensor& s = *stats_rows;
    if (s.scalar_type() != at::kFloat || s.dim() != 2 || s.size(0) != B || s.size(1) != 4 || !s.is_contiguous() || s.device() != dev)
      fail_value("stats_rows must be a contiguous fp32 [B, 4] tensor on the same device");
    r.la.stats_rows = s.data_ptr<float>();
  }
  r.la.clip_range = clip; r.la.adv_clip_max = amax; r.la.kl_coeff = klc; r.la.denom = denom;
  r.la.accumulate = accumulate ? 1 : 0;
  return r;
}

}  // namespace

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
  m.doc() = "compiled binding of include/mixgrpo_b200.h's hot entry points (same C ABI as the ctypes layer)";
  m.def("abi_version", []() { return mixgrpo_abi_version(); });
  m.def("autograd_backward_launches", []() { return g_autograd_backward_launches; });

  m.def("fused_step",
        [](int family, const Tensor& v, const Tensor& x, uint64_t coefs_addr, int src, const OptT& noise, const OptT& x_next, const OptT& m1, const OptT& m2,
           int order, bool sde_solver, const OptT& out_x_next, bool want_x0, bool want_mean, bool want_logp, bool rnd, const OptT& out_logp,
           const OptT& out_x0, int early, bool has_philox, uint64_t ph_seed, uint64_t ph_offset, uint64_t ph_state, const OptT& decode_out, double divisor,
           double shift, bool from_x0, bool reciprocal, uint64_t defer_ptr, int64_t defer_bytes, const OptT& seed_out) {
          StepOut o = fused_step_impl(family, v, x, coefs_at(coefs_addr), src, noise, x_next, m1, m2, order, sde_solver, out_x_next, want_x0, want_mean,
                                      want_logp && !defer_ptr, rnd, out_logp, out_x0, early, has_philox, ph_seed, ph_offset, ph_state, decode_out, divisor,
                                      shift, from_x0, reciprocal, defer_ptr, defer_bytes, seed_out);
          return py::make_tuple(opt(o.x_next), opt(o.x0), opt(o.logp), opt(o.mean));
        });

  m.def("logprob_backward", [](int family, const Tensor& v, const Tensor& x, const Tensor& x_next, const Tensor& grad_logp, uint64_t coefs_addr, bool rnd,
                               const OptT& out) { return logprob_backward_impl(family, v, x, x_next, grad_logp, coefs_at(coefs_addr), rnd, out); });

  // differentiable stored-transition log-prob: returns (logp, x0, mean-or-empty)
  m.def("transition_logprob", [](const Tensor& v, const Tensor& x, const Tensor& x_next, uint64_t coefs_addr, int64_t family, bool rnd, bool want_mean) {
    auto out = TransitionLogProb::apply(v, x, x_next, static_cast<int64_t>(coefs_addr), family, rnd, want_mean);
    return py::make_tuple(out[0], out[1], out[2]);
  });

  m.def("policy_forward", [](int family, Tensor v, Tensor x, Tensor x_next, uint64_t coefs_addr, const Tensor& old_logp, const Tensor& advantages,
                             double clip, double amax, double klc, double denom, const OptT& stats_rows, bool rnd, const OptT& out_logp, bool accumulate,
                             bool early_loads) {
    require_cuda(v, "model_output"); require_cuda(x, "latents"); require_cuda(x_next, "prev_sample");
    const int vd = dtype_code(v, "model_output");
    if (!v.is_contiguous()) v = v.contiguous();
    const auto f32 = v.options().dtype(at::kFloat);
    if (v.size(0) == 0) return out_logp ? *out_logp : at::empty({0}, f32);
    const int64_t B = v.size(0), n = v.numel() / B;
    int64_t x_bs, in_bs;
    std::tie(x, x_bs) = rows(as_f32(x), "latents");
    std::tie(x_next, in_bs) = rows(as_f32(x_next), "prev_sample");
    LossArgsIn L = loss_args(old_logp, advantages, stats_rows, clip, amax, klc, denom, B, v.device(), accumulate);
    Tensor logp = out_logp ? *out_logp : at::empty({B}, f32);
    c10::cuda::OptionalCUDAGuard guard(v.device());
    cudaStream_t st = c10::cuda::getCurrentCUDAStream(v.device().index()).stream();
    Tensor ws = workspace(v.device(), B, n, st);
    check_rc(mixgrpo_policy_fwd(family, v.data_ptr(), vd, x.data_ptr<float>(), x_bs, x_next.data_ptr<float>(), in_bs, logp.data_ptr<float>(), ws.data_ptr(),
                                ws.numel(), B, n, &coefs_at(coefs_addr), &L.la, make_flags(rnd, vd, early_loads ? 2 : 0), st),
             "policy_fwd");
    return logp;
  });

  m.def("policy_backward", [](int family, Tensor v, Tensor x, Tensor x_next, const Tensor& new_logp, uint64_t coefs_addr, const Tensor& old_logp,
                              const Tensor& advantages, double clip, double amax, double klc, double denom, bool rnd, bool early_loads) {
    require_cuda(v, "model_output"); require_cuda(x, "latents"); require_cuda(x_next, "prev_sample"); require_cuda(new_logp, "new_log_probs");
    const int vd = dtype_code(v, "model_output");
    if (!v.is_contiguous()) v = v.contiguous();
    if (v.size(0) == 0) return at::empty_like(v);
    const int64_t B = v.size(0), n = v.numel() / B;
    int64_t x_bs, in_bs;
    std::tie(x, x_bs) = rows(as_f32(x), "latents");
    std::tie(x_next, in_bs) = rows(as_f32(x_next), "prev_sample");
    Tensor nl = as_f32(new_logp.detach()).contiguous().view({-1});
    LossArgsIn L = loss_args(old_logp, advantages, c10::nullopt, clip, amax, klc, denom, B, v.device(), true);
    Tensor grad_v = at::empty_like(v);
    c10::cuda::OptionalCUDAGuard guard(v.device());
    cudaStream_t st = c10::cuda::getCurrentCUDAStream(v.device().index()).stream();
    check_rc(mixgrpo_policy_bwd(family, v.data_ptr(), vd, x.data_ptr<float>(), x_bs, x_next.data_ptr<float>(), in_bs, nl.data_ptr<float>(), &L.la,
                                grad_v.data_ptr(), B, n, &coefs_at(coefs_addr), make_flags(rnd, vd, early_loads ? 2 : 0), st),
             "policy_bwd");
    return grad_v;
  });
}
