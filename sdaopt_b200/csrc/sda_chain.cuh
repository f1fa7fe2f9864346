// sda_chain.cuh -- one annealing chain on one chain group (G lanes): K2 (visit) + K3 (objective) +
// K4 (accept and bookkeeping) + K8 (re-anneal) + K9 (suite monitor), fused.
//
// Restates MarkovChain.run (_sda.py:192-234), EnergyState.reset (:135-166) and the suite wrapper
// Algo._funcwrapped (benchmark/bench.py:292-321) for a group.  Scalars are group-uniform (every
// lane of the group holds the same value); vectors live in the group's shared-memory rows,
// coordinate k owned by lane k mod G.  All barriers are group-masked (sda_device.cuh).
#pragma once

#include "sda_device.cuh"
#include "sda_objectives.cuh"

namespace sda {

// Kernel arguments (device pointers unless noted)
struct DeviceArgs {
    // problem
    int functor, dim;
    const double *lower, *upper, *params;
    // config
    long long maxiter;
    double initial_temp, qv, qa, maxfun, temperature_restart;
    int pure_sa;
    long long max_calls;
    double fglob_tol;  // fglob + tol (NaN disables hits, bench.py:314)
    long long trace_len;
    const double *temp_table, *sigmax_table;  // device copies of the K1 tables
    long long table_len;
    // inputs
    long long n_chains;
    const double *x0;
    const unsigned long long *seeds;
    const double *tape;
    long long tape_len;
    // outputs
    SdaResult out;
    // scratch
    double *xmin;                   // [n_chains * dim]
    double *ls_work;                // [n_slots * ls_stride]
    long long ls_stride;            // doubles per warp slot
    unsigned long long *work_counter;
    int ls_maxiter;                 // clamp(6*dim, 100, 1000)  _sda.py:291-296
    int ls_memo;                    // finite-difference gradients of separable objectives from memoised terms
    int group;                      // lanes per chain (power of two <= 32)
    // Large D: the chain rows (x_cur, x_vis, sv: 3 D doubles per chain) do not fit shared memory at a
    // useful occupancy (D = 1000: 24 KB per chain + a 32 KB box per CTA left 3-4 warps per SM), so they
    // live in this per-warp-slot region of the workspace instead, read and written with lane-striped
    // (coalesced) accesses that the L1/L2 serve; the box stays in shared memory, once per CTA.
    double *rows;                   // [n_slots * (32/group) * 3 * dim] or NULL: rows in shared memory
    // phased execution (annealing launch / local-search launch alternate, sda_kernels.cu)
    int phased, round, steps_per_round;
    struct ChainRec *recs;          // [n_chains] suspended chain state
    double *rec_x;                  // [n_chains * dim] x_cur of suspended chains
    int *ready_q;                   // [2 * n_chains] chains to resume (double buffered by round parity)
    int *ls_q;                      // [n_chains] chains waiting for a local search
    long long *qctl;                // queue counters, see QCtl
};

enum QCtl : int { kReadyHead = 0, kReadyCount = 1, kLsHead = 2, kLsCount = 3, kNextCount = 4 };

// the search box in shared memory: lower, upper, upper - lower and 1 / (upper - lower)
struct Box {
    const double *lo, *up, *range, *inv_range;
};

struct ChainState {
    double e_cur, ebest, emin;
    long long not_improved_idx, not_improved_max_idx;
    long long ncall, ncall_ls, n_ls;
    long long mon_calls, ncall_hit;
    double f_hit;
    long long trace_pos;
    int state_improved, stop, hit, status, have_best;
};

// state of a chain parked between an annealing launch and a local-search launch
struct ChainRec {
    ChainState st;
    long long iter, step_i;
    int err, ls_req;
};

// the callable handed to sda(): objective + suite monitor (bench.py:292-321)
template <bool INLINE_EVAL>
__device__ __forceinline__ double call_func(const DeviceArgs &a, ChainState &st, const double *x,
                                            const Group &g)
{
    // salt: the call number, used only by the objectives that draw random numbers inside fun()
    const unsigned long long salt = (unsigned long long)st.ncall;
    const double res = INLINE_EVAL ? eval_objective(a.functor, x, a.dim, a.params, g, salt)
                                   : eval_objective_ool(a.functor, x, a.dim, a.params, g.G,
                                                        g.gid * g.G + g.gl, salt);
    if (a.max_calls > 0 && !st.stop) {
        st.mon_calls += 1;                              // bench.py:300
        if (st.mon_calls >= a.max_calls) {              // bench.py:301-304
            st.stop = 1;
        } else if (res <= a.fglob_tol) {                // bench.py:314-320
            st.hit = 1;
            st.ncall_hit = st.mon_calls;
            st.f_hit = res;
            st.stop = 1;
        }
    }
    return res;
}

// ObjectiveFunWrapper.func_wrapper                                                :321-323
template <bool IN_LS>
__device__ __forceinline__ double func_wrapper(const DeviceArgs &a, ChainState &st,
                                               const double *x, const Group &g)
{
    st.ncall += 1;
    if (IN_LS) st.ncall_ls += 1;
    return call_func<!IN_LS>(a, st, x, g);
}

__device__ __forceinline__ void copy_vec(double *dst, const double *src, int dim, const Group &g)
{
    for (int k = g.gl; k < dim; k += g.G) dst[k] = src[k];
}

template <bool TAPE>
struct Rng;
template <>
struct Rng<true> { TapeRng t; };
template <>
struct Rng<false> { PhiloxRng p; };

// lower + random_sample(D) * (upper - lower)                                      :137-138, :157-158
template <bool TAPE>
__device__ __forceinline__ void draw_location(Rng<TAPE> &rng, double *x_cur, const Box &bx, int dim,
                                              uint32_t attempt, const Group &g)
{
    if constexpr (TAPE) {
        for (int k = 0; k < dim; ++k) {
            const double u = rng.t.next();
            if ((k & (g.G - 1)) == g.gl)
                x_cur[k] = __dadd_rn(bx.lo[k], __dmul_rn(u, __dsub_rn(bx.up[k], bx.lo[k])));
        }
    } else {
        for (int k = g.gl; k < dim; k += g.G) {
            double u, u2;
            rng.p.pair(kInit, (uint32_t)k, attempt, u, u2);
            x_cur[k] = __dadd_rn(bx.lo[k], __dmul_rn(u, __dsub_rn(bx.up[k], bx.lo[k])));
        }
    }
    group_sync(g);
}

// EnergyState.reset                                                               :135-166
template <bool TAPE>
__device__ __noinline__ int energy_reset(const DeviceArgs &a, ChainState &st, Rng<TAPE> &rng,
                                         double *x_cur, const Box &bx, const double *x0,
                                         double *xbest, const Group &g)
{
    const int dim = a.dim;
    uint32_t attempt = 0;
    if (x0 == nullptr) draw_location<TAPE>(rng, x_cur, bx, dim, attempt++, g);
    else { copy_vec(x_cur, x0, dim, g); group_sync(g); }
    int reinit = 0;
    for (;;) {
        bool done = false;
        st.e_cur = call_func<false>(a, st, x_cur, g);   // owf.func: not counted      :144
        if (st.stop) {
            // the suite monitor ended the run at its very first evaluation (the reference leaves
            // sda() through the monitor's exception here): report that point, not stale memory
            if (!st.have_best) {
                st.ebest = st.e_cur;
                copy_vec(xbest, x_cur, dim, g);
                st.have_best = 1;
            }
            return 0;
        }
        if (st.e_cur >= kBigValue || isnan(st.e_cur)) {
            if (reinit >= kMaxReinit) return SDA_CHAIN_REINIT_FAILED;
            draw_location<TAPE>(rng, x_cur, bx, dim, attempt++, g);
            reinit += 1;
        } else {
            done = true;
        }
        if (!st.have_best) {                            // :163-165
            st.ebest = st.e_cur;
            copy_vec(xbest, x_cur, dim, g);
            st.have_best = 1;
        }
        if (done) break;
    }
    return 0;
}

__device__ __forceinline__ void trace(const DeviceArgs &a, ChainState &st, long long chain,
                                      double e, int dec, const Group &g)
{
    if (a.out.trace_e != nullptr && st.trace_pos < a.trace_len && g.gl == 0) {
        a.out.trace_e[chain * a.trace_len + st.trace_pos] = e;
        a.out.trace_d[chain * a.trace_len + st.trace_pos] = (uint8_t)dec;
    }
    st.trace_pos += 1;
}

// K4: generalized Metropolis acceptance + best / stagnation bookkeeping            :204-233
// x_vis holds the proposal; `single` >= 0 names the only coordinate that differs from x_cur.
// The acceptance uniform r is supplied by `draw_r` only on the non-improving branch (:216).  With
// the counter-based stream it is additionally skipped when pqa == 0: r <= 0 holds only for
// r == 0.0 exactly (probability 2^-53), which the Philox protocol (and the oracle's Philox mode)
// define as a rejection; the tape replay always consumes the uniform like the reference.
template <bool TAPE, class DrawR>
__device__ __forceinline__ void accept_step(const DeviceArgs &a, ChainState &st, long long chain,
                                            double e, DrawR draw_r, double t_step, long long j,
                                            int single, double *x_cur, double *x_vis,
                                            double *xbest, double *xmin, const Group &g)
{
    const int dim = a.dim;
    if (e < st.e_cur) {
        int dec = SDA_DEC_IMPROVE;
        st.e_cur = e;
        if (single >= 0) { if (g.gl == 0) x_cur[single] = x_vis[single]; }
        else copy_vec(x_cur, x_vis, dim, g);
        if (e < st.ebest) {
            st.ebest = e;
            copy_vec(xbest, x_vis, dim, g);
            st.state_improved = 1;
            st.not_improved_idx = 0;
            dec = SDA_DEC_NEW_BEST;
        }
        group_sync(g);
        trace(a, st, chain, e, dec, g);
    } else {
        int dec = SDA_DEC_REJECT;
        const double pqa_temp =
            __dadd_rn(__ddiv_rn(__dmul_rn(a.qa - 1.0, __dsub_rn(e, st.e_cur)), t_step), 1.);
        double pqa;
        bool accept;
        if (pqa_temp < 0.) {
            pqa = 0.;
            if constexpr (TAPE) accept = draw_r() <= pqa;
            else accept = false;
        } else {
            pqa = exp(__ddiv_rn(log(pqa_temp), 1. - a.qa));
            accept = draw_r() <= pqa;
        }
        if (accept) {
            st.e_cur = e;
            if (single >= 0) { if (g.gl == 0) x_cur[single] = x_vis[single]; }
            else copy_vec(x_cur, x_vis, dim, g);
            group_sync(g);
            copy_vec(xmin, x_cur, dim, g);              // self.xmin = np.copy(current_location)
            dec = SDA_DEC_METRO_ACCEPT;
        } else if (single >= 0) {
            if (g.gl == 0) x_vis[single] = x_cur[single];   // undo the one-coordinate proposal
            group_sync(g);
        }
        if (st.not_improved_idx >= st.not_improved_max_idx) {
            if (j == 0 || st.e_cur < st.emin) {
                st.emin = st.e_cur;
                copy_vec(xmin, x_cur, dim, g);
            }
        }
        trace(a, st, chain, e, dec, g);
    }
}

// MarkovChain.run                                                                 :192-234
// VISIT_ILP: coordinates per lane and trip in the visit transform (1: smaller loop body, the
// default everywhere; 2: two interleaved chains, kept for experiments)
template <bool TAPE, int VISIT_ILP>
__device__ __forceinline__ void markov_run(const DeviceArgs &a, ChainState &st, Rng<TAPE> &rng,
                                           long long chain, long long step, double temperature,
                                           double sigmax, double *x_cur, double *x_vis, double *sv,
                                           const Box &bx, double *xbest, double *xmin,
                                           const Group &g)
{
    const int dim = a.dim;
    const int G = g.G;
    const double t_step = __ddiv_rn(temperature, (double)(step + 1));
    const VisitConst vc = { sigmax, a.qv - 1.0, 3.0 - a.qv, (a.qv - 1.0) / (3.0 - a.qv) };
    st.not_improved_idx += 1;
    for (long long j = 0; j < 2LL * dim; ++j) {
        if (j == 0) st.state_improved = 0;
        if (step == 0 && j == 0) st.state_improved = 1;
        int single = -1;
        // ---- K2: proposal                                                       :40-73
        if (j < dim) {
            if constexpr (TAPE) {
                for (int k = 0; k < dim; ++k) {
                    const double v = visit_sample_tape(rng.t, vc);
                    if ((k & (G - 1)) == g.gl) sv[k] = v;
                }
                const double upper_sample = rng.t.next();
                const double lower_sample = rng.t.next();
                group_sync(g);
                for (int k = g.gl; k < dim; k += G) {
                    double v = sv[k];
                    if (v > kTailLimit) v = __dmul_rn(kTailLimit, upper_sample);
                    else if (v < -kTailLimit) v = __dmul_rn(-kTailLimit, lower_sample);
                    x_vis[k] = wrap_coord(__dadd_rn(v, x_cur[k]), bx.lo[k], bx.range[k], bx.inv_range[k]);
                }
            } else {
                rng.p.j = (uint32_t)j;
                // the two clip uniforms are "always drawn" in the reference (:46-47); with a
                // counter-based stream drawing is addressing, so they are computed lazily
                bool have_clip = false;
                double upper_sample = 0.0, lower_sample = 0.0;
                // accepted polar pairs of this lane's coordinates first (flattened rejection loop,
                // scratch: sv <- X, x_vis <- Y; both rows are rewritten below by the same lane)
                polar_fill_philox<false>(rng.p, g.gl, G, dim, sv, x_vis);
                // one coordinate per lane and trip (VISIT_ILP == 1): with the constant-table math the
                // smaller loop body beats two interleaved chains in the 80-register instantiations
                // (382 -> 316 ms pure SA)
                for (int k = g.gl; VISIT_ILP == 1 && k < dim; k += G) {
                    const double xg0 = sv[k], yg0 = x_vis[k];
                    const double s0 = __dadd_rn(__dmul_rn(xg0, xg0), __dmul_rn(yg0, yg0));
                    double v0 = visit_from_polar_fast(xg0, yg0, s0, vc);
                    if (!have_clip && fabs(v0) > kTailLimit) {
                        rng.p.pair(kClip, 0, 0, upper_sample, lower_sample);
                        have_clip = true;
                    }
                    if (v0 > kTailLimit) v0 = __dmul_rn(kTailLimit, upper_sample);
                    else if (v0 < -kTailLimit) v0 = __dmul_rn(-kTailLimit, lower_sample);
                    x_vis[k] = wrap_coord(__dadd_rn(v0, x_cur[k]), bx.lo[k], bx.range[k], bx.inv_range[k]);
                }
                for (int k0 = g.gl; VISIT_ILP == 2 && k0 < dim; k0 += 2 * G) {
                    const int k1 = k0 + G;
                    const bool two = k1 < dim;
                    const double xg0 = sv[k0], yg0 = x_vis[k0];
                    const double xg1 = two ? sv[k1] : 0.5, yg1 = two ? x_vis[k1] : 0.5;
                    const double s0 = __dadd_rn(__dmul_rn(xg0, xg0), __dmul_rn(yg0, yg0));
                    const double s1 = __dadd_rn(__dmul_rn(xg1, xg1), __dmul_rn(yg1, yg1));
                    double v0 = visit_from_polar_fast(xg0, yg0, s0, vc);
                    double v1 = visit_from_polar_fast(xg1, yg1, s1, vc);
                    if (!have_clip && (fabs(v0) > kTailLimit || (two && fabs(v1) > kTailLimit))) {
                        rng.p.pair(kClip, 0, 0, upper_sample, lower_sample);
                        have_clip = true;
                    }
                    if (v0 > kTailLimit) v0 = __dmul_rn(kTailLimit, upper_sample);
                    else if (v0 < -kTailLimit) v0 = __dmul_rn(-kTailLimit, lower_sample);
                    if (v1 > kTailLimit) v1 = __dmul_rn(kTailLimit, upper_sample);
                    else if (v1 < -kTailLimit) v1 = __dmul_rn(-kTailLimit, lower_sample);
                    x_vis[k0] = wrap_coord(__dadd_rn(v0, x_cur[k0]), bx.lo[k0], bx.range[k0], bx.inv_range[k0]);
                    if (two)
                        x_vis[k1] = wrap_coord(__dadd_rn(v1, x_cur[k1]), bx.lo[k1], bx.range[k1], bx.inv_range[k1]);
                }
            }
            group_sync(g);
        } else {
            const int index = (int)(j - dim);
            if (j == dim) {
                // entering the one-coordinate phase: x_vis := x_cur, and (Philox) draw the D
                // state-independent one-coordinate jumps of this step in parallel
                if constexpr (!TAPE) {
                    polar_fill_philox<true>(rng.p, g.gl, G, dim, sv, x_vis);
                    for (int k = g.gl; k < dim; k += G) {
                        rng.p.j = (uint32_t)(dim + k);
                        const double xg = sv[k], yg = x_vis[k];
                        const double s = __dadd_rn(__dmul_rn(xg, xg), __dmul_rn(yg, yg));
                        double v = visit_from_polar_fast(xg, yg, s, vc);
                        if (v > kTailLimit || v < -kTailLimit) {
                            double u, u2;
                            rng.p.pair(kClip, 0, 0, u, u2);
                            v = (v > 0.0) ? __dmul_rn(kTailLimit, u) : __dmul_rn(-kTailLimit, u);
                        }
                        sv[k] = v;
                    }
                }
                copy_vec(x_vis, x_cur, dim, g);     // after the fill: x_vis was its scratch row
                group_sync(g);
            }
            double visit;
            if constexpr (TAPE) {
                visit = visit_sample_tape(rng.t, vc);
                if (visit > kTailLimit) visit = __dmul_rn(kTailLimit, rng.t.next());
                else if (visit < -kTailLimit) visit = __dmul_rn(-kTailLimit, rng.t.next());
            } else {
                rng.p.j = (uint32_t)j;
                visit = sv[index];
            }
            if (g.gl == 0)
                x_vis[index] = wrap_coord(__dadd_rn(visit, x_cur[index]), bx.lo[index],
                                          bx.range[index], bx.inv_range[index]);
            group_sync(g);
            single = index;
        }
        if constexpr (TAPE) { if (rng.t.exhausted) return; }
        // ---- K3: objective                                                      :203
        const double e = func_wrapper<false>(a, st, x_vis, g);
        if (st.stop) return;
        // ---- K4: acceptance.  The uniform is drawn only when e >= E_cur          :216
        auto draw_r = [&]() -> double {
            if constexpr (TAPE) { return rng.t.next(); }
            else { double r, r2; rng.p.pair(kAccept, 0, 0, r, r2); return r; }
        };
        accept_step<TAPE>(a, st, chain, e, draw_r, t_step, j, single, x_cur, x_vis, xbest, xmin, g);
        if constexpr (TAPE) { if (rng.t.exhausted) return; }
    }
}

}  // namespace sda
