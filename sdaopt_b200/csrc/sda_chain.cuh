// sda_chain.cuh -- one annealing chain on one warp: K2 (visit) + K3 (objective) + K4 (accept and
// bookkeeping) + K5 (local-search policy) + K8 (re-anneal) + K9 (suite monitor), fused.
//
// Restates SDARunner.search (_sda.py:393-418), MarkovChain.run (:192-234),
// MarkovChain.local_search (:236-273), EnergyState.reset (:135-166) and the suite wrapper
// Algo._funcwrapped (benchmark/bench.py:292-321) for a warp.  All scalars are warp-uniform.
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
};

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

// the callable handed to sda(): objective + suite monitor (bench.py:292-321)
template <bool INLINE_EVAL>
__device__ __forceinline__ double call_func(const DeviceArgs &a, ChainState &st, const double *x,
                                            int lane)
{
    const double res = INLINE_EVAL ? eval_objective(a.functor, x, a.dim, a.params, lane)
                                   : eval_objective_ool(a.functor, x, a.dim, a.params, lane);
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
                                               const double *x, int lane)
{
    st.ncall += 1;
    if (IN_LS) st.ncall_ls += 1;
    return call_func<!IN_LS>(a, st, x, lane);
}

__device__ __forceinline__ void copy_vec(double *dst, const double *src, int dim, int lane)
{
    for (int k = lane; k < dim; k += SDA_WARP) dst[k] = src[k];
}

template <bool TAPE>
struct Rng;
template <>
struct Rng<true> { TapeRng t; };
template <>
struct Rng<false> { PhiloxRng p; };

// lower + random_sample(D) * (upper - lower)                                      :137-138, :157-158
template <bool TAPE>
__device__ __forceinline__ void draw_location(Rng<TAPE> &rng, double *x_cur, const double *lo,
                                              const double *up, int dim, uint32_t attempt, int lane)
{
    if constexpr (TAPE) {
        for (int k = 0; k < dim; ++k) {
            const double u = rng.t.next();
            if ((k & 31) == lane) x_cur[k] = __dadd_rn(lo[k], __dmul_rn(u, __dsub_rn(up[k], lo[k])));
        }
    } else {
        for (int k = lane; k < dim; k += SDA_WARP) {
            double u, u2;
            rng.p.pair(kInit, (uint32_t)k, attempt, u, u2);
            x_cur[k] = __dadd_rn(lo[k], __dmul_rn(u, __dsub_rn(up[k], lo[k])));
        }
    }
    __syncwarp();
}

// EnergyState.reset                                                               :135-166
template <bool TAPE>
__device__ __noinline__ int energy_reset(const DeviceArgs &a, ChainState &st, Rng<TAPE> &rng,
                                            double *x_cur, const double *lo, const double *up,
                                            const double *x0, double *xbest, int lane)
{
    const int dim = a.dim;
    uint32_t attempt = 0;
    if (x0 == nullptr) draw_location<TAPE>(rng, x_cur, lo, up, dim, attempt++, lane);
    else { copy_vec(x_cur, x0, dim, lane); __syncwarp(); }
    int reinit = 0;
    for (;;) {
        bool done = false;
        st.e_cur = call_func<false>(a, st, x_cur, lane); // owf.func: not counted      :144
        if (st.stop) return 0;
        if (st.e_cur >= kBigValue || isnan(st.e_cur)) {
            if (reinit >= kMaxReinit) return SDA_CHAIN_REINIT_FAILED;
            draw_location<TAPE>(rng, x_cur, lo, up, dim, attempt++, lane);
            reinit += 1;
        } else {
            done = true;
        }
        if (!st.have_best) {                            // :163-165
            st.ebest = st.e_cur;
            copy_vec(xbest, x_cur, dim, lane);
            st.have_best = 1;
        }
        if (done) break;
    }
    return 0;
}

__device__ __forceinline__ void trace(const DeviceArgs &a, ChainState &st, long long chain,
                                      double e, int dec, int lane)
{
    if (a.out.trace_e != nullptr && st.trace_pos < a.trace_len && lane == 0) {
        a.out.trace_e[chain * a.trace_len + st.trace_pos] = e;
        a.out.trace_d[chain * a.trace_len + st.trace_pos] = (uint8_t)dec;
    }
    st.trace_pos += 1;
}

// K4: generalized Metropolis acceptance + best / stagnation bookkeeping            :204-233
// x_vis holds the proposal; `single` >= 0 names the only coordinate that differs from x_cur.
__device__ __forceinline__ void accept_step(const DeviceArgs &a, ChainState &st, long long chain,
                                            double e, double r, double t_step, long long j,
                                            int single, double *x_cur, double *x_vis,
                                            double *xbest, double *xmin, int lane)
{
    const int dim = a.dim;
    if (e < st.e_cur) {
        int dec = SDA_DEC_IMPROVE;
        st.e_cur = e;
        if (single >= 0) { if (lane == 0) x_cur[single] = x_vis[single]; }
        else copy_vec(x_cur, x_vis, dim, lane);
        if (e < st.ebest) {
            st.ebest = e;
            copy_vec(xbest, x_vis, dim, lane);
            st.state_improved = 1;
            st.not_improved_idx = 0;
            dec = SDA_DEC_NEW_BEST;
        }
        __syncwarp();
        trace(a, st, chain, e, dec, lane);
    } else {
        int dec = SDA_DEC_REJECT;
        const double pqa_temp =
            __dadd_rn(__ddiv_rn(__dmul_rn(a.qa - 1.0, __dsub_rn(e, st.e_cur)), t_step), 1.);
        double pqa;
        if (pqa_temp < 0.) pqa = 0.;
        else pqa = exp(__ddiv_rn(log(pqa_temp), 1. - a.qa));
        if (r <= pqa) {
            st.e_cur = e;
            if (single >= 0) { if (lane == 0) x_cur[single] = x_vis[single]; }
            else copy_vec(x_cur, x_vis, dim, lane);
            __syncwarp();
            copy_vec(xmin, x_cur, dim, lane);           // self.xmin = np.copy(current_location)
            dec = SDA_DEC_METRO_ACCEPT;
        } else if (single >= 0) {
            if (lane == 0) x_vis[single] = x_cur[single];   // undo the one-coordinate proposal
            __syncwarp();
        }
        if (st.not_improved_idx >= st.not_improved_max_idx) {
            if (j == 0 || st.e_cur < st.emin) {
                st.emin = st.e_cur;
                copy_vec(xmin, x_cur, dim, lane);
            }
        }
        trace(a, st, chain, e, dec, lane);
    }
}

// MarkovChain.run                                                                 :192-234
template <bool TAPE>
__device__ __forceinline__ void markov_run(const DeviceArgs &a, ChainState &st, Rng<TAPE> &rng,
                                           long long chain, long long step, double temperature,
                                           double sigmax, double *x_cur, double *x_vis, double *sv,
                                           const Box &bx, double *xbest, double *xmin, int lane)
{
    const int dim = a.dim;
    const double t_step = __ddiv_rn(temperature, (double)(step + 1));
    const VisitConst vc = { sigmax, a.qv - 1.0, 3.0 - a.qv };
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
                    if ((k & 31) == lane) sv[k] = v;
                }
                const double upper_sample = rng.t.next();
                const double lower_sample = rng.t.next();
                __syncwarp();
                for (int k = lane; k < dim; k += SDA_WARP) {
                    double v = sv[k];
                    if (v > kTailLimit) v = __dmul_rn(kTailLimit, upper_sample);
                    else if (v < -kTailLimit) v = __dmul_rn(-kTailLimit, lower_sample);
                    x_vis[k] = wrap_coord(__dadd_rn(v, x_cur[k]), bx.lo[k], bx.range[k], bx.inv_range[k]);
                }
            } else {
                rng.p.j = (uint32_t)j;
                double upper_sample, lower_sample;
                rng.p.pair(kClip, 0, 0, upper_sample, lower_sample);
                // two coordinates per lane and trip: the two log/sqrt/exp/div chains are
                // independent, which doubles the ILP of the FP64 pipe at 4 warps per SMSP
                for (int k0 = lane; k0 < dim; k0 += 2 * SDA_WARP) {
                    const int k1 = k0 + SDA_WARP;
                    const bool two = k1 < dim;
                    double xg0, yg0, s0, xg1 = 0.5, yg1 = 0.5, s1 = 0.5;
                    polar_pair_philox(rng.p, (uint32_t)k0, xg0, yg0, s0);
                    if (two) polar_pair_philox(rng.p, (uint32_t)k1, xg1, yg1, s1);
                    double v0 = visit_from_polar(xg0, yg0, s0, vc);
                    double v1 = visit_from_polar(xg1, yg1, s1, vc);
                    if (v0 > kTailLimit) v0 = __dmul_rn(kTailLimit, upper_sample);
                    else if (v0 < -kTailLimit) v0 = __dmul_rn(-kTailLimit, lower_sample);
                    if (v1 > kTailLimit) v1 = __dmul_rn(kTailLimit, upper_sample);
                    else if (v1 < -kTailLimit) v1 = __dmul_rn(-kTailLimit, lower_sample);
                    x_vis[k0] = wrap_coord(__dadd_rn(v0, x_cur[k0]), bx.lo[k0], bx.range[k0], bx.inv_range[k0]);
                    if (two)
                        x_vis[k1] = wrap_coord(__dadd_rn(v1, x_cur[k1]), bx.lo[k1], bx.range[k1], bx.inv_range[k1]);
                }
            }
            __syncwarp();
        } else {
            const int index = (int)(j - dim);
            if (j == dim) {
                // entering the one-coordinate phase: x_vis := x_cur, and (Philox) draw the D
                // state-independent one-coordinate jumps of this step in parallel
                copy_vec(x_vis, x_cur, dim, lane);
                if constexpr (!TAPE) {
                    for (int k = lane; k < dim; k += SDA_WARP) {
                        rng.p.j = (uint32_t)(dim + k);
                        double v = visit_sample_philox(rng.p, (uint32_t)k, vc);
                        if (v > kTailLimit || v < -kTailLimit) {
                            double u, u2;
                            rng.p.pair(kClip, 0, 0, u, u2);
                            v = (v > 0.0) ? __dmul_rn(kTailLimit, u) : __dmul_rn(-kTailLimit, u);
                        }
                        sv[k] = v;
                    }
                }
                __syncwarp();
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
            if (lane == 0)
                x_vis[index] = wrap_coord(__dadd_rn(visit, x_cur[index]), bx.lo[index],
                                          bx.range[index], bx.inv_range[index]);
            __syncwarp();
            single = index;
        }
        if constexpr (TAPE) { if (rng.t.exhausted) return; }
        // ---- K3: objective                                                      :203
        const double e = func_wrapper<false>(a, st, x_vis, lane);
        if (st.stop) return;
        // ---- K4: acceptance.  The uniform is drawn only when e >= E_cur          :216
        double r = 0.0;
        if (!(e < st.e_cur)) {
            if constexpr (TAPE) { r = rng.t.next(); if (rng.t.exhausted) return; }
            else { double r2; rng.p.pair(kAccept, 0, 0, r, r2); }
        }
        accept_step(a, st, chain, e, r, t_step, j, single, x_cur, x_vis, xbest, xmin, lane);
    }
}

}  // namespace sda
