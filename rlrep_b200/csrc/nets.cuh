// Network building blocks shared by the SPEDER-SAC and Diff-SR-SAC handles:
//   Trunk     -- util.mlp / spedersac mlp (utils/util.py:85-96, spedersac_agent.py:67-76): Linear, ELU, ..., Linear
//   RffCritic -- the sin -> ELU -> linear twin critic (spedersac_agent.py:21-50, diffsrsac_agent.py:40-90)
// Both are thin launch sequences over the GEMM runner; parameters live in the caller's ParamGroup.
#pragma once
#include "agent.cuh"

namespace rlrep {

struct Trunk {
  std::vector<LinearSlot> l;
  std::vector<int> width;  // output width of each Linear
  int in = 0;

  // hidden_depth == 0 is a single Linear(in, out); otherwise Linear(in, h), ELU, (Linear(h, h), ELU) x (depth - 1),
  // Linear(h, out).  Sequential indices of the Linears are 0, 2, 4, ... (state_dict names `<prefix>.<index>`).
  void plan(ParamGroup& g, const std::string& prefix, int in_dim, int hidden, int out, int depth) {
    in = in_dim;
    int cur = in_dim;
    for (int i = 0; i <= depth; ++i) {
      const int o = i == depth ? out : hidden;
      RLREP_CHECK(o % 32 == 0 || i == depth, "trunk hidden widths must be multiples of 32");
      l.push_back(add_linear(g, prefix + "." + std::to_string(2 * i), o, cur));
      width.push_back(o);
      cur = o;
    }
  }
  int n() const { return (int)l.size(); }
  int out() const { return width.back(); }
};

// Hidden activations (and their gradients) of one use of a trunk on up to `rows` rows.
struct TrunkActs {
  std::vector<float*> h, dh;
  void want(DeviceArena& a, const Trunk& t, int rows, bool with_grad) {
    h.assign(t.n() - 1, nullptr);
    dh.assign(t.n() - 1, nullptr);
    for (int i = 0; i + 1 < t.n(); ++i) {
      a.want(&h[i], (size_t)rows * t.width[i]);
      if (with_grad) a.want(&dh[i], (size_t)rows * t.width[i]);
    }
  }
};

inline void trunk_forward(GemmRunner& g, cudaStream_t s, int rows, const Trunk& t, const ParamGroup& pg, bool target,
                          Mat x, Mat x2, int k1, TrunkActs& a, float* out, int ld_out) {
  Mat cur = x;
  for (int i = 0; i < t.n(); ++i) {
    const Linear li = t.l[i].view(pg, target);
    const bool last = i + 1 == t.n();
    float* y = last ? out : a.h[i];
    const int ldy = last ? ld_out : t.width[i];
    linear_fwd(g, s, rows, cur, li, last ? ACT_NONE : ACT_ELU, y, ldy, i == 0 ? x2 : Mat(), i == 0 ? k1 : 0);
    cur = Mat{y, ldy};
  }
}

// dout = gradient w.r.t. the trunk output.  wgrad: also leaves dW in the group's gradient arena and appends the bias
// jobs (one batched column reduction per network is launched by the caller).  dx != nullptr: gradient w.r.t. input
// columns [col0, col0 + n_cols).
inline void trunk_backward(GemmRunner& g, cudaStream_t s, int rows, const Trunk& t, const ParamGroup& pg, bool target,
                           Mat dout, Mat x, TrunkActs& a, bool wgrad, std::vector<ColJob>* bias_jobs, float* dx,
                           int ld_dx, int col0, int n_cols) {
  Mat dy = dout;
  for (int i = t.n() - 1; i >= 0; --i) {
    const Linear li = t.l[i].view(pg, target);
    const Mat xin = i == 0 ? x : Mat{a.h[i - 1], t.width[i - 1]};
    if (wgrad) {
      linear_wgrad(g, s, rows, dy, xin, li, Mat(), 0, false);
      bias_jobs->push_back(bias_job(rows, dy, li));
    }
    if (i > 0) {
      linear_dgrad(g, s, rows, dy, li, DACT_ELU_OUT, xin, a.dh[i - 1], t.width[i - 1]);
      dy = Mat{a.dh[i - 1], t.width[i - 1]};
    } else if (dx != nullptr) {
      linear_dgrad(g, s, rows, dy, li, DACT_NONE, Mat(), dx, ld_dx, col0, n_cols);
    }
  }
}

inline void launch_bias_jobs(const std::vector<ColJob>& jobs, cudaStream_t s) {
  for (size_t i = 0; i < jobs.size(); i += kMaxColJobs)
    launch_colreduce_multi(jobs.data() + i, (int)std::min<size_t>(kMaxColJobs, jobs.size() - i), s);
}

// Twin critic on features z [B, D]:  q1 = l3(elu(l2(sin(l1 z)))),  q2 = l6(elu(l5(sin(l4 z)))).
// l1 | l4 are stored stacked so the two random-feature layers are ONE [2H, D] GEMM; the pre-activation is kept for
// the cosine in the backward pass.
struct RffCritic {
  LinearSlot c14, c2, c5, c3, c6;
  int H = 0, D = 0, B = 0;
  // one activation set per concurrent use (target critic on s', live critic on s)
  float *pre[2] = {nullptr, nullptr}, *sn[2] = {nullptr, nullptr}, *hid2[2] = {nullptr, nullptr}, *q[2] = {nullptr, nullptr};
  float *dhid2 = nullptr, *dsn = nullptr;

  void plan(ParamGroup& g, DeviceArena& a, int D_, int H_, int B_) {
    D = D_; H = H_; B = B_;
    RLREP_CHECK(H % 32 == 0 && D % 32 == 0, "critic hidden_dim and feature_dim must be multiples of 32");
    c14 = add_stacked(g, "critic.l1", H, "critic.l4", H, D);
    c2 = add_linear(g, "critic.l2", H, H);
    c5 = add_linear(g, "critic.l5", H, H);
    c3 = add_linear(g, "critic.l3", 1, H, false);
    c6 = add_linear(g, "critic.l6", 1, H, false);
    for (int i = 0; i < 2; ++i) {
      a.want(&pre[i], (size_t)B * 2 * H);
      a.want(&sn[i], (size_t)B * 2 * H);
      a.want(&hid2[i], (size_t)2 * B * H);
      a.want(&q[i], 2 * B);
    }
    a.want(&dhid2, (size_t)2 * B * H);
    a.want(&dsn, (size_t)B * 2 * H);
  }

  // q1 -> q[slot][0:B], q2 -> q[slot][B:2B]
  void forward(GemmRunner& g, cudaStream_t s, const ParamGroup& pg, bool target, int slot, const float* z) {
    const Linear l14 = c14.view(pg, target), l2 = c2.view(pg, target), l5 = c5.view(pg, target);
    const Linear l3 = c3.view(pg, target), l6 = c6.view(pg, target);
    linear_fwd(g, s, B, Mat{z, D}, l14, ACT_SIN, sn[slot], 2 * H, Mat(), 0, pre[slot]);
    linear_fwd(g, s, B, Mat{sn[slot], 2 * H}, l2, ACT_ELU, hid2[slot], H);
    linear_fwd(g, s, B, Mat{sn[slot] + H, 2 * H}, l5, ACT_ELU, hid2[slot] + (size_t)B * H, H);
    launch_rowdot_pair(RowDotJob{hid2[slot], l3.W, l3.b, q[slot], H, H},
                       RowDotJob{hid2[slot] + (size_t)B * H, l6.W, l6.b, q[slot] + B, H, H}, B, s);
  }

  // (dq1 | dq2) [2B] -> parameter gradients (wgrad) and / or the gradient w.r.t. z (dz [B, D]).
  void backward(GemmRunner& g, cudaStream_t s, const ParamGroup& pg, int slot, const float* z, const float* dq,
                bool wgrad, float* dz) {
    const Linear l14 = c14.view(pg), l2 = c2.view(pg), l5 = c5.view(pg), l3 = c3.view(pg), l6 = c6.view(pg);
    float* dha = dhid2;
    float* dhb = dhid2 + (size_t)B * H;
    const float* ha = hid2[slot];
    const float* hb = hid2[slot] + (size_t)B * H;
    launch_outer_dact(dq, l3.W, B, H, ha, H, DACT_ELU_OUT, dha, H, s);
    launch_outer_dact(dq + B, l6.W, B, H, hb, H, DACT_ELU_OUT, dhb, H, s);
    if (wgrad) {
      linear_wgrad(g, s, B, Mat{dha, H}, Mat{sn[slot], 2 * H}, l2, Mat(), 0, false);
      linear_wgrad(g, s, B, Mat{dhb, H}, Mat{sn[slot] + H, 2 * H}, l5, Mat(), 0, false);
    }
    linear_dgrad(g, s, B, Mat{dha, H}, l2, DACT_COS_PRE, Mat{pre[slot], 2 * H}, dsn, 2 * H);
    linear_dgrad(g, s, B, Mat{dhb, H}, l5, DACT_COS_PRE, Mat{pre[slot] + H, 2 * H}, dsn + H, 2 * H);
    if (wgrad) {
      linear_wgrad(g, s, B, Mat{dsn, 2 * H}, Mat{z, D}, l14, Mat(), 0, false);
      const ColJob jobs[7] = {ColJob{ha, dq, l3.dW, H, B, H},     ColJob{dq, nullptr, l3.db, 1, B, 1},
                              ColJob{hb, dq + B, l6.dW, H, B, H}, ColJob{dq + B, nullptr, l6.db, 1, B, 1},
                              bias_job(B, Mat{dha, H}, l2),       bias_job(B, Mat{dhb, H}, l5),
                              bias_job(B, Mat{dsn, 2 * H}, l14)};
      launch_colreduce_multi(jobs, 7, s);
    }
    if (dz != nullptr) linear_dgrad(g, s, B, Mat{dsn, 2 * H}, l14, DACT_NONE, Mat(), dz, D);
  }
};

}  // namespace rlrep
