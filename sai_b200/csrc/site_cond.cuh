// Per-site conditions with the reference's float64 semantics.
//
// Follows compute_matching_loci (sai/stats/stat_utils.py:114-166):
//   freq      = num / (called * ploidy)            IEEE double division (:49-52)
//   valid     = every freq finite and in [0,1]     (:121-130)
//   match_y   = AND_k op_k(src_freq_k, y_k)        (:141-144)
//   match_1my = AND_k op_k(src_freq_k, 1 - y_k)    only without ancestral alleles,
//               with `1 - y_k` evaluated by the host in Python floats (:148-152)
//   inverted  = match_1my & valid -> ref/tgt freq := 1 - freq   (:156-160)
//   cond      = valid & (match_y | match_1my) & ref_freq < w    (:166)
// and UStatistic adds tgt_freq > x (sai/stats/u_statistic.py:92); QStatistic
// keeps tgt_freq of the cond sites (sai/stats/q_statistic.py:92).
// Every arithmetic step uses the explicitly rounded intrinsics so that no FMA
// contraction or fast-math flag can change a decision.
#pragma once
#include "common.cuh"

namespace sai {

struct SiteFlags {
  bool u;
  bool q;
  double q_tgt_freq;
};

__device__ __forceinline__ bool cmp_op(int op, double f, double y) {
  switch (op) {
    case SAI_OP_EQ: return f == y;
    case SAI_OP_LT: return f < y;
    case SAI_OP_GT: return f > y;
    case SAI_OP_LE: return f <= y;
    default: return f >= y;
  }
}

// One condition block; returns cond and the (possibly inverted) tgt frequency.
__device__ __forceinline__ bool eval_cond(const sai_cond& c, int n_src, bool anc, double fr,
                                          double ft, const double* fs, double& tgt_out) {
  bool my = true, mf = true;
#pragma unroll
  for (int k = 0; k < SAI_MAX_SRC; ++k) {
    if (k < n_src) {
      my = my && cmp_op(c.op[k], fs[k], c.y[k]);
      mf = mf && cmp_op(c.op[k], fs[k], c.one_minus_y[k]);
    }
  }
  const bool inv = !anc && mf;
  const bool match = my || inv;
  const double r = inv ? __dsub_rn(1.0, fr) : fr;
  tgt_out = inv ? __dsub_rn(1.0, ft) : ft;
  return match && (r < c.w);
}

template <typename NumFn, typename CalledFn>
__device__ __forceinline__ SiteFlags eval_site(const sai_job& J, const sai_layout& lay,
                                               NumFn num_of, CalledFn called_of) {
  SiteFlags out{false, false, 0.0};
  const int nr = num_of(J.ref_pop), dr = called_of(J.ref_pop) * lay.pop[J.ref_pop].ploidy;
  const int nt = num_of(J.tgt_pop), dt = called_of(J.tgt_pop) * lay.pop[J.tgt_pop].ploidy;
  // freq in [0,1] and finite  <=>  den > 0 and 0 <= num <= den  (num is never negative)
  bool valid = dr > 0 && nr <= dr && dt > 0 && nt <= dt;
  double fs[SAI_MAX_SRC];
#pragma unroll
  for (int k = 0; k < SAI_MAX_SRC; ++k) {
    if (k < J.n_src) {
      const int sp = J.src_pop[k];
      const int ns = num_of(sp), ds = called_of(sp) * lay.pop[sp].ploidy;
      valid = valid && ds > 0 && ns <= ds;
      fs[k] = ds > 0 ? __ddiv_rn((double)ns, (double)ds) : 0.0;
    } else {
      fs[k] = 0.0;
    }
  }
  if (!valid) return out;
  const double fr = __ddiv_rn((double)nr, (double)dr);
  const double ft = __ddiv_rn((double)nt, (double)dt);
  const bool anc = J.anc_allele_available != 0;
  if (J.u.enabled) {
    double t;
    const bool c = eval_cond(J.u, J.n_src, anc, fr, ft, fs, t);
    out.u = c && (t > J.x);
  }
  if (J.q.enabled) {
    double t;
    const bool c = eval_cond(J.q, J.n_src, anc, fr, ft, fs, t);
    out.q = c;
    out.q_tgt_freq = t;
  }
  return out;
}

}  // namespace sai
