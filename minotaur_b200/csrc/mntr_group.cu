// mntr_group.cu -- node batches over several GPUs of one box from ONE process (include/mntr_gpu.h, "group" calls).
//
// A group is one engine context per device.  The problem is replicated (the CSR and the tapes are small next to the
// boxes); a node batch is split contiguously over the members and every member's share runs on its own host thread
// and its own device stream -- boxes are independent, so there is no collective (node-batch mode).
// Host C++ only: the kernels are the single-context ones.
#include "../../include/mntr_gpu.h"

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <new>
#include <string>
#include <thread>
#include <vector>

struct mntr_gpu_group {
  std::vector<mntr_gpu_ctx *> members;
  char err[512] = {0};
};

namespace {

int group_fail(mntr_gpu_group *g, int code, const char *what, const mntr_gpu_ctx *c)
{
  if (g) snprintf(g->err, sizeof(g->err), "%s: %s", what, c ? mntr_gpu_last_error(c) : "bad argument");
  return code;
}

// runs f(member index) on one host thread per member; returns the first non-zero code
template <class F>
int for_each_member(mntr_gpu_group *g, const char *what, F f)
{
  const size_t k = g->members.size();
  std::vector<int> rc(k, 0);
  std::vector<std::thread> th;
  for (size_t i = 1; i < k; ++i) th.emplace_back([&, i] { rc[i] = f((int)i); });
  rc[0] = f(0);
  for (auto &t : th) t.join();
  for (size_t i = 0; i < k; ++i)
    if (rc[i] != 0) return group_fail(g, rc[i], what, g->members[i]);
  return MNTR_OK;
}

}  // namespace

extern "C" {

int mntr_gpu_group_create(int32_t n_devices, const int32_t *devices, mntr_gpu_group **out)
{
  if (!out) return MNTR_E_ARG;
  *out = nullptr;
  if (n_devices <= 0 || !devices) return MNTR_E_ARG;
  for (int32_t i = 0; i < n_devices; ++i)
    for (int32_t k = 0; k < i; ++k)
      if (devices[i] == devices[k]) return MNTR_E_ARG;      // one member per device
  mntr_gpu_group *g = new (std::nothrow) mntr_gpu_group();
  if (!g) return MNTR_E_NOMEM;
  for (int32_t i = 0; i < n_devices; ++i) {
    mntr_gpu_ctx *c = nullptr;
    const int rc = mntr_gpu_create(devices[i], &c);
    if (rc != MNTR_OK) { mntr_gpu_group_destroy(g); return rc; }      // no CPU fallback: every device must be usable
    g->members.push_back(c);
  }
  *out = g;
  return MNTR_OK;
}

void mntr_gpu_group_destroy(mntr_gpu_group *g)
{
  if (!g) return;
  for (mntr_gpu_ctx *c : g->members) mntr_gpu_destroy(c);
  delete g;
}

int32_t mntr_gpu_group_size(const mntr_gpu_group *g) { return g ? (int32_t)g->members.size() : 0; }

mntr_gpu_ctx *mntr_gpu_group_member(mntr_gpu_group *g, int32_t i)
{
  return (g && i >= 0 && i < (int32_t)g->members.size()) ? g->members[(size_t)i] : nullptr;
}

const char *mntr_gpu_group_last_error(const mntr_gpu_group *g) { return g ? g->err : "null group"; }

int mntr_gpu_group_load_linear(mntr_gpu_group *g, int32_t m, int32_t n, const int32_t *row_ptr, const int32_t *col,
                               const double *val, const double *row_lb, const double *row_ub, const uint8_t *var_type,
                               const uint8_t *row_active)
{
  if (!g) return MNTR_E_ARG;
  return for_each_member(g, "group_load_linear", [&](int i) {
    return mntr_gpu_load_linear(g->members[(size_t)i], m, n, row_ptr, col, val, row_lb, row_ub, var_type, row_active); });
}

int mntr_gpu_group_load_cgraph(mntr_gpu_group *g, int32_t n_cons, const int32_t *tape_ptr, const uint8_t *op,
                               const int32_t *arg0, const int32_t *arg1, const double *cnst, const int32_t *child,
                               const int32_t *lin_ptr, const int32_t *lin_col, const double *lin_val, const double *c_lb,
                               const double *c_ub)
{
  if (!g) return MNTR_E_ARG;
  return for_each_member(g, "group_load_cgraph", [&](int i) {
    return mntr_gpu_load_cgraph(g->members[(size_t)i], n_cons, tape_ptr, op, arg0, arg1, cnst, child, lin_ptr, lin_col, lin_val,
                                c_lb, c_ub); });
}

int mntr_gpu_group_load_quad(mntr_gpu_group *g, int32_t n_quad, const int32_t *q_ptr, const int32_t *v1, const int32_t *v2,
                             const double *coef, const int32_t *lin_ptr, const int32_t *lin_col, const double *lin_val,
                             const double *q_lb, const double *q_ub)
{
  if (!g) return MNTR_E_ARG;
  return for_each_member(g, "group_load_quad", [&](int i) {
    return mntr_gpu_load_quad(g->members[(size_t)i], n_quad, q_ptr, v1, v2, coef, lin_ptr, lin_col, lin_val, q_lb, q_ub); });
}

int mntr_gpu_group_set_cutoff(mntr_gpu_group *g, int32_t k, const int32_t *col, const double *val, double rhs)
{
  if (!g) return MNTR_E_ARG;
  return for_each_member(g, "group_set_cutoff", [&](int i) { return mntr_gpu_set_cutoff(g->members[(size_t)i], k, col, val, rhs); });
}

int mntr_gpu_group_set_incumbent(mntr_gpu_group *g, double best_value)
{
  if (!g) return MNTR_E_ARG;
  return for_each_member(g, "group_set_incumbent", [&](int i) { return mntr_gpu_set_incumbent(g->members[(size_t)i], best_value); });
}

int mntr_gpu_group_tighten_nodes(mntr_gpu_group *g, int32_t n_boxes, const double *root_lb, const double *root_ub,
                                 const int64_t *delta_ptr, const int32_t *delta_var, const uint8_t *delta_is_upper,
                                 const double *delta_val, const mntr_gpu_options *opts, int32_t *verdict, int32_t *rounds,
                                 int64_t *mod_ptr, int32_t *mod_var, uint8_t *mod_is_upper, double *mod_val, int64_t mod_cap,
                                 int64_t *n_mods_out)
{
  if (!g || g->members.empty()) return MNTR_E_ARG;
  if (n_boxes <= 0 || !delta_ptr || !mod_ptr || mod_cap < 0) return group_fail(g, MNTR_E_ARG, "group_tighten_nodes", nullptr);
  const int k = (int)std::min<size_t>(g->members.size(), (size_t)((n_boxes + 31) / 32));   // whole tiles of 32 boxes
  // contiguous shares, multiples of 32 boxes (the kernels' tile) except the last
  const int32_t tiles = (n_boxes + 31) / 32;
  std::vector<int32_t> b0((size_t)k + 1, 0);
  for (int i = 0; i <= k; ++i) b0[(size_t)i] = std::min<int64_t>(n_boxes, (int64_t)tiles * i / k * 32);
  struct Share {
    std::vector<int64_t> dptr, mptr;
    std::vector<int32_t> mvar; std::vector<uint8_t> mup; std::vector<double> mval;
    int64_t total = 0;
    int rc = 0;
  };
  std::vector<Share> sh((size_t)k);
  auto run = [&](int i) {
    Share &S = sh[(size_t)i];
    const int32_t lo = b0[(size_t)i], nb = b0[(size_t)i + 1] - lo;
    if (nb <= 0) return;
    const int64_t base = delta_ptr[lo];
    S.dptr.resize((size_t)nb + 1);
    for (int32_t b = 0; b <= nb; ++b) S.dptr[(size_t)b] = delta_ptr[lo + b] - base;
    S.mptr.assign((size_t)nb + 1, 0);
    S.mvar.resize((size_t)std::max<int64_t>(mod_cap, 1)); S.mup.resize(S.mvar.size()); S.mval.resize(S.mvar.size());
    S.rc = mntr_gpu_tighten_nodes(g->members[(size_t)i], nb, root_lb, root_ub, S.dptr.data(),
                                  delta_var ? delta_var + base : nullptr, delta_is_upper ? delta_is_upper + base : nullptr,
                                  delta_val ? delta_val + base : nullptr, opts, verdict ? verdict + lo : nullptr,
                                  rounds ? rounds + lo : nullptr, S.mptr.data(), S.mvar.data(), S.mup.data(), S.mval.data(),
                                  mod_cap, &S.total);
  };
  std::vector<std::thread> th;
  for (int i = 1; i < k; ++i) th.emplace_back(run, i);
  run(0);
  for (auto &t : th) t.join();
  for (int i = 0; i < k; ++i)
    if (sh[(size_t)i].rc != 0) return group_fail(g, sh[(size_t)i].rc, "group_tighten_nodes", g->members[(size_t)i]);
  // stitch the shares together in box order
  int64_t total = 0;
  for (int i = 0; i < k; ++i) total += sh[(size_t)i].total;
  if (n_mods_out) *n_mods_out = total;
  int64_t at = 0;
  for (int i = 0; i < k; ++i) {
    const Share &S = sh[(size_t)i];
    const int32_t lo = b0[(size_t)i], nb = b0[(size_t)i + 1] - lo;
    for (int32_t b = 0; b < nb; ++b) mod_ptr[lo + b] = at + S.mptr[(size_t)b];
    if (total <= mod_cap && S.total > 0) {
      memcpy(mod_var + at, S.mvar.data(), sizeof(int32_t) * (size_t)S.total);
      memcpy(mod_is_upper + at, S.mup.data(), (size_t)S.total);
      memcpy(mod_val + at, S.mval.data(), sizeof(double) * (size_t)S.total);
    }
    at += S.total;
  }
  mod_ptr[n_boxes] = at;
  return MNTR_OK;
}

}  // extern "C"
