// sda_kernels.cu -- kernels and the C ABI of libsdaopt_b200.so (include/sdaopt_b200.h).
//
// One persistent launch per batch: every warp pulls chain indices from a device-side counter and
// runs whole chains (SDARunner.search, _sda.py:393-418) start to finish, including the local
// searches, so heavy-tailed run lengths (suite early exits, L-BFGS-B) balance dynamically and no
// host round trip happens between temperature steps.
#include <type_traits>
#include <cuda_runtime.h>

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "sda_chain.cuh"
#include "sda_lbfgsb.cuh"

namespace sda {

#ifndef SDA_ANNEAL_MIN_BLOCKS
#define SDA_ANNEAL_MIN_BLOCKS 2
#endif

// what a chain group does next (per-group state machine of the persistent kernel)
enum Pending : int { kNone = 0, kDoInit = 10, kReanneal = 11, kStep = 12 };
// local-search request of a group, served by the whole warp
enum LsReq : int { kLsNone = 0, kLsFromBest = 1, kLsFromMin = 2 };

__device__ __forceinline__ long long shfl_ll(long long v, int src)
{
    return __shfl_sync(SDA_FULL_MASK, v, src);
}

// make the chain state of the group led by lane `src` warp-uniform (for the warp-wide L-BFGS-B)
__device__ __forceinline__ ChainState bcast_state(const ChainState &s, int src)
{
    ChainState t;
    t.e_cur = __shfl_sync(SDA_FULL_MASK, s.e_cur, src);
    t.ebest = __shfl_sync(SDA_FULL_MASK, s.ebest, src);
    t.emin = __shfl_sync(SDA_FULL_MASK, s.emin, src);
    t.not_improved_idx = shfl_ll(s.not_improved_idx, src);
    t.not_improved_max_idx = shfl_ll(s.not_improved_max_idx, src);
    t.ncall = shfl_ll(s.ncall, src);
    t.ncall_ls = shfl_ll(s.ncall_ls, src);
    t.n_ls = shfl_ll(s.n_ls, src);
    t.mon_calls = shfl_ll(s.mon_calls, src);
    t.ncall_hit = shfl_ll(s.ncall_hit, src);
    t.f_hit = __shfl_sync(SDA_FULL_MASK, s.f_hit, src);
    t.trace_pos = shfl_ll(s.trace_pos, src);
    t.state_improved = __shfl_sync(SDA_FULL_MASK, s.state_improved, src);
    t.stop = __shfl_sync(SDA_FULL_MASK, s.stop, src);
    t.hit = __shfl_sync(SDA_FULL_MASK, s.hit, src);
    t.status = __shfl_sync(SDA_FULL_MASK, s.status, src);
    t.have_best = __shfl_sync(SDA_FULL_MASK, s.have_best, src);
    return t;
}

// The chain state and the L-BFGS-B work descriptor are handed to the (noinline) minimiser as
// copies so that the annealing loop's own state never has its address taken and stays in registers.
template <class Clock>
__device__ __forceinline__ int call_local_search(const DeviceArgs &a, ChainState &st, double *ls_slot,
                                                 double *ls_smem, const double *x_start, double *xe,
                                                 const double *lo, const double *up, double &e,
                                                 int lane, Clock &clk)
{
    ChainState tmp = st;
    LsWork w;
    w.bind(ls_slot, ls_smem, a.dim);
    const int ok = ofw_local_search(a, tmp, w, x_start, xe, lo, up, e, lane, clk);
    st = tmp;
    return ok;
}
__device__ __forceinline__ int call_local_search(const DeviceArgs &a, ChainState &st, double *ls_slot,
                                                 double *ls_smem, const double *x_start, double *xe,
                                                 const double *lo, const double *up, double &e,
                                                 int lane)
{
    NoClock clk;
    return call_local_search(a, st, ls_slot, ls_smem, x_start, xe, lo, up, e, lane, clk);
}

// MarkovChain.local_search after the minimiser returned (:244-250, :261-273): executed by the lanes
// that own the chain (a group in the persistent kernel, the whole warp in the local-search launch).
// Returns the follow-up request (the stagnation search) or kLsNone.
__device__ __forceinline__ int apply_ls_result(const DeviceArgs &a, ChainState &st, int req, int ok,
                                               double e, const double *wx, double *xbest,
                                               double *xmin, double *x_cur, const Group &g)
{
    const int dim = a.dim;
    int next = kLsNone;
    if (req == kLsFromBest) {
        if (e < st.ebest) {                                          // :244-250
            st.not_improved_idx = 0;
            st.ebest = e;
            copy_vec(xbest, wx, dim, g);
            st.e_cur = e;
            copy_vec(x_cur, wx, dim, g);
        } else if (st.not_improved_idx >= st.not_improved_max_idx) {  // :261-262
            next = kLsFromMin;
        }
    } else {
        if (ok) copy_vec(xmin, wx, dim, g);                          // :265
        st.emin = e;
        st.not_improved_idx = 0;
        st.not_improved_max_idx = dim;
        if (e < st.ebest) {                                          // :269-273
            st.ebest = st.emin;
            copy_vec(xbest, wx, dim, g);
            st.e_cur = e;
            copy_vec(x_cur, wx, dim, g);
        }
    }
    return next;
}

// SDARunner.result :420-428 (+ the suite record, bench.py:184-188)
__device__ __forceinline__ void write_result(const DeviceArgs &a, long long chain, const ChainState &st,
                                             long long iter, int err, long long tape_pos)
{
    a.out.fun[chain] = st.ebest;
    a.out.nit[chain] = iter;
    a.out.ncall[chain] = st.ncall;
    a.out.status[chain] = err;
    if (a.out.ncall_ls) a.out.ncall_ls[chain] = st.ncall_ls;
    if (a.out.n_ls) a.out.n_ls[chain] = st.n_ls;
    if (a.out.hit) a.out.hit[chain] = st.hit;
    if (a.out.ncall_hit) a.out.ncall_hit[chain] = st.ncall_hit;
    if (a.out.f_hit) a.out.f_hit[chain] = st.f_hit;
    if (a.out.tape_used) a.out.tape_used[chain] = tape_pos;
}

template <bool TAPE>
__device__ __forceinline__ int call_energy_reset(const DeviceArgs &a, ChainState &st, Rng<TAPE> &rng,
                                                 double *x_cur, const Box &bx, const double *x0,
                                                 double *xbest, const Group &g)
{
    ChainState tmp = st;
    Rng<TAPE> r2 = rng;
    const int err = energy_reset<TAPE>(a, tmp, r2, x_cur, bx, x0, xbest, g);
    st = tmp;
    rng = r2;
    return err;
}

// shared memory: lower[D] upper[D] range[D] 1/range[D]
//              | per warp: 32/G chains x { x_cur[D] x_vis[D] sv[D] } | per warp: L-BFGS-B matrices
//
// Persistent kernel.  Every chain group runs the state machine of SDARunner.search (_sda.py:393-418)
// for one chain at a time: INIT -> { STEP | REANNEAL }* -> finished -> pull the next chain from the
// device counter.  One pass of the warp loop = at most one (re)initialisation, one Markov step
// (MarkovChain.run, 2*D evaluations) and the local searches that step triggers, for every group of
// the warp; full-mask barriers between the phases keep the groups aligned so that 32/G chains
// share every issued instruction, whatever step, chain or function-call count each one is at.
//
// PHASED = true is the annealing half of the phased execution (see sda_ls_phase_kernel): a group
// that needs a local search parks its chain in global memory, queues it and pulls the next ready
// chain; the local-search code is not part of this instantiation, which therefore holds 3 CTAs
// per SM.
#ifndef SDA_VISIT_ILP_PERSISTENT
#define SDA_VISIT_ILP_PERSISTENT 1   // 2 measured equal (Sutton-Chen N >= 128) or slower (D=50: -7 %)
#endif
#ifndef SDA_GUIDED_MIN
#define SDA_GUIDED_MIN 1
#endif
#ifndef SDA_GUIDED_WAVES
#define SDA_GUIDED_WAVES 1
#endif
#ifndef SDA_PHASED_MIN_BLOCKS
// resident CTAs per SM of the two instantiations without local-search code.  4 -> 64 registers, 32
// warps per SM) wins the first, hot third of a run (same-box A/B on the headline shape at maxiter 100 / 300,
// profiles/r2_ab_occupancy_estrin.log: +7 % / +6 %) but loses it back at low temperature: over the
// whole maxiter = 1000 run 3 is 1.5 % faster (7.26 vs 7.37 s, profiles/r2_ab_split_compile_bisect.log).
#define SDA_PHASED_MIN_BLOCKS 3
#endif
// MODE: kPersistent = annealing with in-kernel local search; kPhasedAnneal = annealing half of the
// phased execution; kPureSa = pure_sa runs (no local-search code either: the same 80-register,
// 3-CTA shape as the phased half, +11 % over running pure SA through the persistent instantiation).
enum KernelMode { kPersistent = 0, kPhasedAnneal = 1, kPureSa = 2 };
// GROWS: the chain rows live in a.rows (global memory, large D) instead of shared memory -- a template
// parameter, not a run-time select, so that the shared-memory instantiation keeps LDS/STS addressing
// (a run-time pointer select made every row access a generic load: -8 % on the headline shape).
#ifndef SDA_GROWS_MIN_BLOCKS
#define SDA_GROWS_MIN_BLOCKS 4   // rows in global memory: occupancy hides their latency (64 registers, 32 warps/SM)
#endif
template <bool TAPE, int MODE, bool GROWS>
__global__ void __launch_bounds__(256, MODE != kPersistent ? (GROWS ? SDA_GROWS_MIN_BLOCKS : SDA_PHASED_MIN_BLOCKS)
                                                           : SDA_ANNEAL_MIN_BLOCKS)
    sda_anneal_kernel(const DeviceArgs a)
{
    constexpr bool PHASED = MODE == kPhasedAnneal;
    constexpr bool NO_LS = MODE != kPersistent;
    extern __shared__ double smem[];
    const int dim = a.dim;
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int warps_per_block = blockDim.x >> 5;
    const int G = a.group;
    const int NG = SDA_WARP / G;
    const Group g = make_group(G, lane);
    double *lo = smem;
    double *up = smem + dim;
    double *rg = smem + 2 * dim;
    double *irg = smem + 3 * dim;
    for (int k = threadIdx.x; k < dim; k += blockDim.x) {
        const double l = a.lower[k], u = a.upper[k];
        lo[k] = l; up[k] = u;
        rg[k] = __dsub_rn(u, l);                       // self.b_range = ub - lb      :35
        irg[k] = __ddiv_rn(1.0, rg[k]);
    }
    __syncthreads();
    const Box bx = { lo, up, rg, irg };
    const long long slot = (long long)blockIdx.x * warps_per_block + warp;
    constexpr bool rows_global = GROWS;
    double *warp_rows;
    if constexpr (GROWS) warp_rows = a.rows + (size_t)slot * NG * 3 * dim;
    else warp_rows = smem + 4 * dim + (size_t)warp * NG * 3 * dim;
    double *x_cur = warp_rows + (size_t)g.gid * 3 * dim;
    double *x_vis = x_cur + dim;
    double *sv = x_vis + dim;
    double *ls_slot = a.ls_work ? a.ls_work + slot * a.ls_stride : nullptr;
    double *ls_smem = NO_LS ? nullptr : smem + 4 * dim + (rows_global ? 0 : (size_t)warps_per_block * NG * 3 * dim) +
                                             (size_t)warp * ls_smem_doubles();
    // a.phased && round > 0: the chains come from the ready queue and resume from their parked state
    // (also used by the non-PHASED instantiation to finish what the bounded rounds left over)
    const bool resume = a.phased && a.round > 0;
    const long long n_ready = resume ? a.qctl[kReadyCount] : a.n_chains;
    const int *ready_cur = a.phased ? a.ready_q + (size_t)(a.round & 1) * a.n_chains : nullptr;
    unsigned long long *pull_counter =
        a.phased ? reinterpret_cast<unsigned long long *>(a.qctl + kReadyHead) : a.work_counter;

    // ---- per-group state -------------------------------------------------------------------
    ChainState st;
    Rng<TAPE> rng;
    long long chain = -1, iter = 0, step_i = 0;
    int pending = kNone, err = 0, steps_done = 0, quota = 0;
    bool have = false, queue_empty = false;
    double *xbest = nullptr, *xmin = nullptr;

    for (;;) {
        // ---- phase 0: idle groups pull the next chain -------------------------------------------
        if (!have && !queue_empty) {
            unsigned long long c = 0;
            if (g.gl == 0) c = atomicAdd(pull_counter, 1ULL);
            c = __shfl_sync(g.mask, c, 0, G);
            if ((long long)c < n_ready) {
                chain = resume ? (long long)ready_cur[c] : (long long)c;
                have = true;
                st.e_cur = st.ebest = st.emin = 0.0;
                st.not_improved_idx = 0; st.not_improved_max_idx = 1000;     // :186-187
                st.ncall = st.ncall_ls = st.n_ls = 0;
                st.mon_calls = 0; st.ncall_hit = a.max_calls;                // bench.py:269
                st.f_hit = __longlong_as_double(0x7ff0000000000000LL);       // bench.py:272
                st.trace_pos = 0;
                st.state_improved = 0; st.stop = 0; st.hit = 0; st.status = 0; st.have_best = 0;
                if constexpr (TAPE) {
                    rng.t.tape = a.tape + chain * a.tape_len;
                    rng.t.pos = 0; rng.t.len = a.tape_len; rng.t.exhausted = false;
                } else {
                    const unsigned long long s = a.seeds ? a.seeds[chain] : (unsigned long long)chain;
                    rng.p.ph.k0 = (uint32_t)s; rng.p.ph.k1 = (uint32_t)(s >> 32);
                    rng.p.iter = 0; rng.p.j = 0xFFFFFFu;
                }
                xbest = a.out.x + chain * dim;
                xmin = a.xmin + chain * dim;
                iter = 0; step_i = 0; err = 0; steps_done = 0;
                if constexpr (PHASED) {
                    // guided scheduling of the launch's tail: once less than one chain per group is
                    // left in the queue the step quota halves (then quarters, ...), so the groups
                    // run out of work within ~1 step of each other instead of ~steps_per_round/2
                    // (ncu: 14 % of the group slots of a launch were idle).  A chain that yields
                    // early simply continues in the next round; results do not depend on it.
                    const long long left = n_ready - (long long)c;
                    const long long groups = (long long)gridDim.x * warps_per_block * NG;
                    quota = a.steps_per_round;
                    for (long long lim = groups * SDA_GUIDED_WAVES; quota > SDA_GUIDED_MIN && left < lim; lim >>= 1) quota >>= 1;
                }
                pending = kDoInit;
                if (resume) {
                    // resume a chain parked by an earlier launch
                    const ChainRec &rec = a.recs[chain];
                    st = rec.st; iter = rec.iter; step_i = rec.step_i; err = rec.err;
                    copy_vec(x_cur, a.rec_x + chain * dim, dim, g);
                    group_sync(g);
                    pending = kStep;
                }
            } else {
                queue_empty = true;
            }
        }
        if (!__any_sync(SDA_FULL_MASK, have)) break;

        // ---- phase 1: SDARunner.__init__ :375-376 / re-anneal :408-410 ----------------------------
        if (have && (pending == kDoInit || pending == kReanneal)) {
            const double *x0 = (pending == kDoInit && a.x0) ? a.x0 + chain * dim : nullptr;
            err = call_energy_reset<TAPE>(a, st, rng, x_cur, bx, x0, xbest, g);
            if (pending == kDoInit && !err && !st.stop) {
                st.emin = st.e_cur;                                      // :176-177
                copy_vec(xmin, x_cur, dim, g);
            }
            pending = kStep;
        }
        __syncwarp();

        // ---- phase 2: the loop header of SDARunner.search :397-410 -------------------------------
        bool do_step = false, finished = false;
        double temperature = 0.0, sigmax = 0.0;
        if (have) {
            bool tape_out = false;
            if constexpr (TAPE) tape_out = rng.t.exhausted;
            if (err || st.stop || tape_out || a.maxiter <= 0) {
                finished = true;
            } else {
                const long long ti = step_i < a.table_len ? step_i : a.table_len - 1;
                temperature = a.temp_table[ti];
                sigmax = a.sigmax_table[ti];
                iter += 1;
                if constexpr (!TAPE) { rng.p.iter = (uint32_t)iter; rng.p.j = 0xFFFFFFu; }
                if (iter == a.maxiter) finished = true;                                // :404-406
                else if (temperature < a.temperature_restart) { pending = kReanneal; step_i = 0; } // :408-410
                else do_step = true;
            }
        }

        // ---- phase 3: MarkovChain.run :412 ---------------------------------------------------------
        if (do_step)
            markov_run<TAPE, (MODE == kPersistent ? SDA_VISIT_ILP_PERSISTENT : 1)>(a, st, rng, chain, step_i, temperature, sigmax, x_cur, x_vis, sv, bx,
                             xbest, xmin, g);
        __syncwarp();

        // ---- phase 4: after the step :413-418, and the local-search policy :236-273 -----------------
        int ls_req = kLsNone;
        if (do_step) {
            bool tape_out = false;
            if constexpr (TAPE) tape_out = rng.t.exhausted;
            if (st.stop || tape_out) {
                finished = true;
            } else if ((double)st.ncall >= a.maxfun) {
                step_i = 0;                                 // `break`: the while loop restarts at i = 0
            } else if (MODE != kPureSa && !a.pure_sa) {
                if (st.state_improved) ls_req = kLsFromBest;                                   // :241
                else if (st.not_improved_idx >= st.not_improved_max_idx) ls_req = kLsFromMin;  // :261
                step_i += 1;
            } else {
                step_i += 1;
            }
        }
        if constexpr (PHASED) {
            // park the chain: for the local-search launch if it posted a request, or straight for
            // the next annealing launch when it used up its steps of this round (bounded rounds
            // keep the launches balanced: no group is left running one long chain alone)
            if (do_step) steps_done += 1;
            const bool yield = have && !finished && ls_req == kLsNone && steps_done >= quota;
            if (have && (ls_req != kLsNone || yield)) {
                if (g.gl == 0) {
                    ChainRec &rec = a.recs[chain];
                    rec.st = st; rec.iter = iter; rec.step_i = step_i; rec.err = err; rec.ls_req = ls_req;
                    if (yield) {
                        const unsigned long long pos =
                            atomicAdd(reinterpret_cast<unsigned long long *>(a.qctl + kNextCount), 1ULL);
                        a.ready_q[(size_t)((a.round + 1) & 1) * a.n_chains + pos] = (int)chain;
                    } else {
                        const unsigned long long pos =
                            atomicAdd(reinterpret_cast<unsigned long long *>(a.qctl + kLsCount), 1ULL);
                        a.ls_q[pos] = (int)chain;
                    }
                }
                copy_vec(a.rec_x + chain * dim, x_cur, dim, g);
                ls_req = kLsNone;
                have = false;
                pending = kNone;
            }
        }
        // two serving rounds: a search from xbest that does not improve may be followed by the
        // stagnation search from xmin (:251-273)
        for (int round = 0; round < 2 && !NO_LS; ++round) {
            if (!__any_sync(SDA_FULL_MASK, ls_req != kLsNone)) break;
            for (int gi = 0; gi < NG; ++gi) {
                const int src = gi * G;
                const int req = __shfl_sync(SDA_FULL_MASK, ls_req, src);
                if (req == kLsNone) continue;
                // the whole warp minimises for group gi: warp-uniform copies of its state
                ChainState ts = bcast_state(st, src);
                const long long tchain = shfl_ll(chain, src);
                double *t_xbest = a.out.x + tchain * dim;
                double *t_xmin = a.xmin + tchain * dim;
                double *t_xe = warp_rows + (size_t)gi * 3 * dim + dim;   // that group's x_vis row
                double e;
                const int ok = call_local_search(a, ts, ls_slot, ls_smem,
                                                 req == kLsFromBest ? t_xbest : t_xmin, t_xe, lo, up,
                                                 e, lane);
                const double *wx = ls_slot + 2LL * kLsM * dim + 6LL * dim;   // LsWork::x
                if (g.gid == gi) {
                    // counters and monitor state come back from the warp-uniform copy
                    st.ncall = ts.ncall; st.ncall_ls = ts.ncall_ls; st.n_ls = ts.n_ls;
                    st.mon_calls = ts.mon_calls; st.ncall_hit = ts.ncall_hit; st.f_hit = ts.f_hit;
                    st.hit = ts.hit; st.stop = ts.stop;
                    if (st.stop) {
                        ls_req = kLsNone;
                        finished = true;
                    } else {
                        ls_req = apply_ls_result(a, st, req, ok, e, wx, xbest, xmin, x_cur, g);
                    }
                    if (!finished && (double)st.ncall >= a.maxfun) step_i = 0;        // :417-418
                }
                __syncwarp();
            }
        }

        // ---- phase 5: SDARunner.result :420-428 ----------------------------------------------------
        if (have && finished) {
            if constexpr (TAPE) { if (rng.t.exhausted) err |= SDA_CHAIN_TAPE_EXHAUSTED; }
            if (g.gl == 0) {
                long long tape_pos = 0;
                if constexpr (TAPE) tape_pos = rng.t.pos;
                write_result(a, chain, st, iter, err, tape_pos);
            }
            have = false;
            pending = kNone;
        }
        __syncwarp();
    }
}

// The local-search half of the phased execution: every warp pulls parked chains from the queue the
// annealing launch filled, minimises (both requests of MarkovChain.local_search if the first does
// not improve), applies the result to the parked state and queues the chain for the next
// annealing launch -- or writes its result if the suite monitor stopped it.
// shared memory: lower[D] upper[D] | per warp: xe[D] + L-BFGS-B matrices
#ifndef SDA_LS_MIN_BLOCKS
#define SDA_LS_MIN_BLOCKS 4
#endif
// SYNC: the warps of the CTA (SDA_LS_SYNC_WARPS of them, 16 warps per SM in all) advance their searches in
// lock step on the phase clock of sda_lbfgsb.cuh so that an SM fetches one phase's code at a time.
#ifndef SDA_LS_SYNC_WARPS
#define SDA_LS_SYNC_WARPS 8     // two CTAs (two clocks) per SM; 16 and 4 measured 7 % and 4.5 % slower
#endif
#ifndef SDA_LS_SYNC_BLOCKS
#define SDA_LS_SYNC_BLOCKS 2    // resident CTAs (clocks) per SM of the synchronous launch
#endif
template <bool SYNC, bool GROWS>
__global__ void __launch_bounds__(SYNC ? SDA_LS_SYNC_WARPS * 32 : 128, SYNC ? SDA_LS_SYNC_BLOCKS : SDA_LS_MIN_BLOCKS)
    sda_ls_phase_kernel(const DeviceArgs a)
{
    extern __shared__ double smem[];
    __shared__ int done_warps;
    if (threadIdx.x == 0) done_warps = 0;
    typename std::conditional<SYNC, PhaseClock, NoClock>::type clk;
    if constexpr (SYNC) { clk.cur = 0; clk.nthreads = (int)blockDim.x; }
    const int dim = a.dim;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    double *lo = smem, *up = smem + dim;
    for (int k = threadIdx.x; k < dim; k += blockDim.x) { lo[k] = a.lower[k]; up[k] = a.upper[k]; }
    __syncthreads();
    const long long slot = (long long)blockIdx.x * wpb + warp;
    // evaluation buffer of the warp: shared memory, or (large D) a row of its slot of a.rows
    double *xe, *ls_smem;
    if constexpr (GROWS) {
        xe = a.rows + (size_t)slot * (SDA_WARP / a.group) * 3 * dim;
        ls_smem = smem + 2 * dim + (size_t)warp * ls_smem_doubles();
    } else {
        xe = smem + 2 * dim + (size_t)warp * (dim + ls_smem_doubles());
        ls_smem = xe + dim;
    }
    double *ls_slot = a.ls_work + slot * a.ls_stride;
    const long long n_ls = a.qctl[kLsCount];
    int *ready_next = a.ready_q + (size_t)((a.round + 1) & 1) * a.n_chains;
    const Group w32 = make_group(SDA_WARP, lane);
    for (;;) {
        unsigned long long c = 0;
        if (lane == 0) c = atomicAdd(reinterpret_cast<unsigned long long *>(a.qctl + kLsHead), 1ULL);
        c = __shfl_sync(SDA_FULL_MASK, c, 0);
        if ((long long)c >= n_ls) break;
        const long long chain = a.ls_q[c];
        // (SYNC) requests are popped in any phase; the search itself starts at the next kPhEval
        const ChainRec rec = a.recs[chain];
        ChainState st = rec.st;
        long long step_i = rec.step_i;
        int req = rec.ls_req;
        bool finished = false;
        double *xbest = a.out.x + chain * dim;
        double *xmin = a.xmin + chain * dim;
        double *x_cur = a.rec_x + chain * dim;
        for (int r = 0; r < 2 && req != kLsNone; ++r) {
            double e;
            const int ok = call_local_search(a, st, ls_slot, ls_smem, req == kLsFromBest ? xbest : xmin,
                                             xe, lo, up, e, lane, clk);
            const double *wx = ls_slot + 2LL * kLsM * dim + 6LL * dim;   // LsWork::x
            if (st.stop) { finished = true; req = kLsNone; }
            else req = apply_ls_result(a, st, req, ok, e, wx, xbest, xmin, x_cur, w32);
            if (!finished && (double)st.ncall >= a.maxfun) step_i = 0;                 // :417-418
            __syncwarp();
        }
        if (lane == 0) {
            if (finished) {
                write_result(a, chain, st, rec.iter, rec.err, 0);
            } else {
                ChainRec &out = a.recs[chain];
                out.st = st; out.step_i = step_i; out.ls_req = kLsNone;
                const unsigned long long pos =
                    atomicAdd(reinterpret_cast<unsigned long long *>(a.qctl + kNextCount), 1ULL);
                ready_next[pos] = (int)chain;
            }
        }
        __syncwarp();
    }
    if constexpr (SYNC) {
        // out of work: keep the clock ticking for the warps that still search.  Warps sign off only
        // in kPhStep and the count is read only right after the barrier that opens kPhDir, so every
        // warp of the CTA sees the same count and they all leave on the same tick.
        clk.enter(kPhStep);
        if (lane == 0) atomicAdd(&done_warps, 1);
        for (;;) {
            clk.tick();
            if (clk.cur == PhaseClock::read_slot() && *(volatile int *)&done_warps == wpb) break;
        }
    }
}

// queue bookkeeping between the launches of a round (one thread)
__global__ void sda_rotate_kernel(long long *qctl, int after_ls)
{
    if (after_ls) {
        qctl[kReadyCount] = qctl[kNextCount];
        qctl[kNextCount] = 0;
        qctl[kReadyHead] = 0;
        qctl[kLsCount] = 0;
    } else {
        qctl[kLsHead] = 0;
    }
}

// objective evaluation at n points, one chain group (G lanes, as in the annealing kernel) per point
__global__ void __launch_bounds__(256) sda_eval_kernel(int functor, int dim, const double *params,
                                                       const double *x, long long n, double *f, int G)
{
    const int lane = threadIdx.x & 31;
    const Group g = make_group(G, lane);
    const long long gidx = ((long long)blockIdx.x * blockDim.x + threadIdx.x) / G;
    const long long ngroups = ((long long)gridDim.x * blockDim.x) / G;
    const long long rounds = (n + ngroups - 1) / ngroups;
    for (long long r = 0; r < rounds; ++r) {
        const long long p = gidx + r * ngroups;
        if (p < n) {
            const double v = eval_objective(functor, x + p * dim, dim, params, g);
            if (g.gl == 0) f[p] = v;
        }
    }
}

// VisitingDistribution.visiting on a tape (unit-test hook)                           :40-73
__global__ void sda_visiting_kernel(DeviceArgs a, const double *x, long long step, double sigmax,
                                    double *x_visit, long long *tape_used)
{
    extern __shared__ double smem[];
    const int dim = a.dim, lane = threadIdx.x & 31;
    double *sv = smem;
    TapeRng rng = { a.tape, 0, a.tape_len, false };
    const VisitConst vc = { sigmax, a.qv - 1.0, 3.0 - a.qv, (a.qv - 1.0) / (3.0 - a.qv) };
    if (step < dim) {
        for (int k = 0; k < dim; ++k) {
            const double v = visit_sample_tape(rng, vc);
            if ((k & 31) == lane) sv[k] = v;
        }
        const double us = rng.next(), ls = rng.next();
        __syncwarp();
        for (int k = lane; k < dim; k += SDA_WARP) {
            double v = sv[k];
            if (v > kTailLimit) v = __dmul_rn(kTailLimit, us);
            else if (v < -kTailLimit) v = __dmul_rn(-kTailLimit, ls);
            { const double rg = __dsub_rn(a.upper[k], a.lower[k]); x_visit[k] = wrap_coord(__dadd_rn(v, x[k]), a.lower[k], rg, __ddiv_rn(1.0, rg)); }
        }
    } else {
        const int index = (int)(step - dim);
        for (int k = lane; k < dim; k += SDA_WARP) x_visit[k] = x[k];
        double visit = visit_sample_tape(rng, vc);
        if (visit > kTailLimit) visit = __dmul_rn(kTailLimit, rng.next());
        else if (visit < -kTailLimit) visit = __dmul_rn(-kTailLimit, rng.next());
        __syncwarp();
        if (lane == 0)
        {
            const double rg = __dsub_rn(a.upper[index], a.lower[index]);
            x_visit[index] = wrap_coord(__dadd_rn(visit, x[index]), a.lower[index], rg, __ddiv_rn(1.0, rg));
        }
    }
    if (lane == 0) *tape_used = rng.exhausted ? -1 : rng.pos;
}

// n successive visit_fn(T) values from a tape (unit-test hook)                       :75-91
__global__ void sda_visit_fn_kernel(double qv, double sigmax, const double *tape, long long tape_len,
                                    double *out, long long n, long long *tape_used)
{
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    TapeRng rng = { tape, 0, tape_len, false };
    const VisitConst vc = { sigmax, qv - 1.0, 3.0 - qv, (qv - 1.0) / (3.0 - qv) };
    for (long long i = 0; i < n; ++i) out[i] = visit_sample_tape(rng, vc);
    *tape_used = rng.exhausted ? -1 : rng.pos;
}

// ObjectiveFunWrapper.local_search for n start points, one warp each (unit-test hook) :347-351
__global__ void __launch_bounds__(128) sda_local_search_kernel(DeviceArgs a, const double *x_in,
                                                              long long n, double *x_out,
                                                              double *f_out, int *success,
                                                              long long *nfev)
{
    extern __shared__ double smem[];
    const int dim = a.dim;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    double *lo = smem, *up = smem + dim;
    for (int k = threadIdx.x; k < dim; k += blockDim.x) { lo[k] = a.lower[k]; up[k] = a.upper[k]; }
    __syncthreads();
    double *xe = smem + 2 * dim + (size_t)warp * dim;
    const long long slot = (long long)blockIdx.x * wpb + warp;
    const long long nslots = (long long)gridDim.x * wpb;
    LsWork w;
    w.bind(a.ls_work + slot * a.ls_stride, smem + (size_t)(2 + wpb) * dim + (size_t)warp * ls_smem_doubles(), dim);
    for (long long p = slot; p < n; p += nslots) {
        ChainState st;
        memset(&st, 0, sizeof(st));
        double e;
        NoClock clk;
        const int ok = ofw_local_search(a, st, w, x_in + p * dim, xe, lo, up, e, lane, clk);
        for (int k = lane; k < dim; k += SDA_WARP) x_out[p * dim + k] = w.x[k];
        if (lane == 0) { f_out[p] = e; success[p] = ok; nfev[p] = st.ncall; }
        __syncwarp();
    }
}

// EnergyState.reset on a tape, one warp (unit-test hook)                            :135-166
// scal: [0] current_energy, [1] ebest (in: the retained best when have_best, out: the best after)
__global__ void sda_energy_reset_kernel(DeviceArgs a, const double *x0, int have_best, double *x_cur_out,
                                        double *xbest_io, double *scal, int *status, long long *tape_used)
{
    extern __shared__ double smem[];
    const int dim = a.dim, lane = threadIdx.x & 31;
    const Group g = make_group(SDA_WARP, lane);
    double *lo = smem, *up = smem + dim, *rg = smem + 2 * dim, *irg = smem + 3 * dim, *x_cur = smem + 4 * dim;
    for (int k = lane; k < dim; k += SDA_WARP) {
        const double l = a.lower[k], u = a.upper[k];
        lo[k] = l; up[k] = u; rg[k] = __dsub_rn(u, l); irg[k] = __ddiv_rn(1.0, rg[k]);
    }
    __syncwarp();
    const Box bx = { lo, up, rg, irg };
    ChainState st;
    memset(&st, 0, sizeof(st));
    st.have_best = have_best;
    st.ebest = scal[1];
    Rng<true> rng;
    rng.t.tape = a.tape; rng.t.pos = 0; rng.t.len = a.tape_len; rng.t.exhausted = false;
    const int err = energy_reset<true>(a, st, rng, x_cur, bx, x0, xbest_io, g);
    __syncwarp();
    for (int k = lane; k < dim; k += SDA_WARP) x_cur_out[k] = x_cur[k];
    if (lane == 0) {
        scal[0] = st.e_cur;
        scal[1] = st.ebest;
        *status = err;
        *tape_used = rng.t.exhausted ? -1 : rng.t.pos;
    }
}

// accuracy probe of sda_math.cuh
__global__ void sda_math_probe_kernel(int kind, const double *x, long long n, double *out)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double v = x[i];
    out[i] = kind == 0 ? cos_2pi(v) : kind == 1 ? log_pos(v) : kind == 2 ? exp_any(v)
           : kind == 3 ? sqrt_normal(v) : div_normal(1.2345678901234567, v);
}

// FP64 FMA throughput probe: 8 independent DFMA chains per thread
__global__ void __launch_bounds__(256) sda_dfma_kernel(double *out, int iters)
{
    double a0 = threadIdx.x * 1e-9, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5,
           a6 = a0 + 6, a7 = a0 + 7;
    const double b = 1.0000001, c = 1e-7;
    for (int i = 0; i < iters; ++i) {
        a0 = fma(a0, b, c); a1 = fma(a1, b, c); a2 = fma(a2, b, c); a3 = fma(a3, b, c);
        a4 = fma(a4, b, c); a5 = fma(a5, b, c); a6 = fma(a6, b, c); a7 = fma(a7, b, c);
    }
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

}  // namespace sda

// =================================================================================================
// host side
// =================================================================================================
using sda::DeviceArgs;

static thread_local char g_cuda_err[256] = "";

static int cuda_fail(cudaError_t e, const char *what)
{
    snprintf(g_cuda_err, sizeof(g_cuda_err), "%s: %s", what, cudaGetErrorString(e));
    return SDA_E_CUDA;
}
#define CK(call)                                                   \
    do {                                                           \
        cudaError_t e_ = (call);                                   \
        if (e_ != cudaSuccess) return cuda_fail(e_, #call);        \
    } while (0)

static const char *kFunctorNames[SDA_FN_COUNT] = {
    "Rastrigin", "ShiftedRastrigin", "Rosenbrock", "Ackley01", "Schwefel01", "Schwefel26",
    "Exponential", "SuttonChen", "Constant", "Sphere", "Griewank",
    "Alpine01", "Brown", "Cigar", "CosineMixture", "Deb01", "DixonPrice", "Easom", "Qing", "Quintic", "Salomon", "Schwefel20", "Schwefel21", "Schwefel22", "Step", "StyblinskiTang", "Trid", "Zacharov", "SixHumpCamel", "Beale", "HimmelBlau", "Matyas", "McCormick", "GoldsteinPrice", "Levy13", "Bohachevsky1", "Bukin06", "EggCrate", "Schaffer01", "DropWave", "ThreeHumpCamel", "Zettl", "Leon", "Cube", "Brent", "Price01", "Bird", "Ackley02", "Ackley03", "Zirilli", "Wavy", "Step2", "Plateau", "Sodp", "Trigonometric02", "XinSheYang02", "SineEnvelope", "Pathological", "Rana", "Adjiman",
    "AMGM", "Alpine02", "BartelsConn", "BiggsExp02", "BiggsExp03", "BiggsExp04", "BiggsExp05",
    "Bohachevsky2", "Bohachevsky3", "BoxBetts", "Branin01", "Branin02", "Bukin02", "Bukin04",
    "CarromTable", "Chichinadze", "Cola", "Colville", "Corana", "CrossInTray", "CrossLegTable",
    "CrownedCross", "Csendes", "Damavandi", "DeVilliersGlasser01", "DeVilliersGlasser02", "Deb03",
    "Decanomial", "Deceptive", "DeckkersAarts", "DeflectedCorrugatedSpring", "Dolan", "Eckerle4",
    "EggHolder", "ElAttarVidyasagarDutta", "Exp2", "FreudensteinRoth", "Gear", "Giunta", "Gulf",
    "Hansen", "Hartmann3", "Hartmann6", "HelicalValley", "HolderTable", "Hosaki", "Infinity",
    "JennrichSampson", "Judge", "Katsuura", "Keane", "Kowalik", "Langermann", "LennardJones",
    "Levy03", "Levy05", "Meyer", "Michalewicz", "MieleCantrell", "Mishra01", "Mishra02", "Mishra03",
    "Mishra04", "Mishra05", "Mishra06", "Mishra07", "Mishra08", "Mishra09", "Mishra10", "Mishra11",
    "MultiModal", "NeedleEye", "NewFunction01", "NewFunction02", "OddSquare", "Parsopoulos",
    "Paviani", "PenHolder", "Penalty01", "Penalty02", "PermFunction01", "PermFunction02", "Pinter",
    "Powell", "PowerSum", "Price02", "Price03", "Price04", "Quadratic", "Ratkowsky01",
    "Ratkowsky02", "Ripple01", "Ripple25", "RosenbrockModified", "RotatedEllipse01",
    "RotatedEllipse02", "Sargan", "Schaffer02", "Schaffer03", "Schaffer04", "Schwefel02",
    "Schwefel04", "Schwefel06", "Schwefel36", "Shekel05", "Shekel07", "Shekel10", "Shubert01",
    "Shubert03", "Shubert04", "StretchedV", "TestTubeHolder", "Thurber", "Treccani", "Trefethen",
    "Trigonometric01", "Tripod", "Ursem01", "Ursem03", "Ursem04", "UrsemWaves",
    "VenterSobiezcczanskiSobieski", "Vincent", "Watson", "WayburnSeader01", "WayburnSeader02",
    "Weierstrass", "Whitley", "Wolfe", "XinSheYang03", "XinSheYang04", "Xor", "YaoLiu04",
    "YaoLiu09", "ZeroSum", "Zimmerman", "Stochastic", "XinSheYang01",
};

extern "C" int sda_abi_version(void) { return SDA_ABI_VERSION; }

extern "C" int sda_has_user_functor(void)
{
#ifdef SDA_USER_FUNCTOR_HEADER
    return 1;
#else
    return 0;
#endif
}

static bool functor_ok(int f)
{
    if (f >= 0 && f < SDA_FN_COUNT) return true;
    return f == SDA_FN_USER && sda_has_user_functor();
}

// Arity of the catalogue classes whose fun() indexes x[0..N-1] unconditionally (their `N` in
// go_benchmark_functions/go_funcs_*.py with change_dimensionality = False); 0 = any dimension.
// The reference raises IndexError (or silently ignores extra coordinates) for a wrong-sized x; here a
// mismatch would read past the chain's shared-memory row, so it is rejected up front.
static const unsigned char kFunctorArity[SDA_FN_COUNT] = {
    0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 2, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 2, 2,
    2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 0, 0, 0, 0, 0, 0, 0, 2, 0, 2,
    0, 0, 2, 2, 3, 4, 5, 2, 2, 3, 2, 2, 2, 2, 2, 2, 17, 4, 4, 2, 2, 2, 0, 2, 4, 5, 0, 2, 0, 2,
    0, 5, 3, 0, 2, 2, 2, 4, 2, 3, 2, 3, 6, 3, 2, 2, 0, 2, 2, 0, 2, 4, 2, 0, 2, 2, 3, 2, 4, 0,
    0, 2, 2, 2, 2, 0, 2, 3, 2, 0, 0, 0, 2, 2, 2, 2, 10, 2, 0, 0, 0, 0, 0, 4, 4, 2, 2, 2, 0, 4,
    3, 2, 2, 2, 2, 2, 0, 2, 2, 2, 0, 0, 2, 2, 4, 4, 4, 0, 0, 0, 0, 2, 7, 2, 2, 0, 2, 2, 2, 2,
    2, 2, 0, 6, 2, 2, 0, 0, 3, 0, 0, 9, 0, 0, 0, 2, 0, 0,
};

extern "C" int sda_functor_arity(int functor)
{
    if (functor < 0 || functor >= SDA_FN_COUNT) return SDA_E_INVALID;
    return kFunctorArity[functor];
}

// functor id known, dimension admissible for it
static bool problem_shape_ok(int functor, int dim)
{
    if (!functor_ok(functor) || dim <= 0) return false;
    if (functor == SDA_FN_USER) return true;
    if (kFunctorArity[functor] && dim != kFunctorArity[functor]) return false;
    if ((functor == SDA_FN_SUTTON_CHEN || functor == SDA_FN_LENNARD_JONES) && dim % 3 != 0) return false;
    return true;
}

// SM count of the current device (cached per device ordinal)
static int device_sms(void)
{
    static int cached[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    if (cached[dev] == 0) {
        int sms = 0;
        if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) return 148;
        cached[dev] = sms;
    }
    return cached[dev];
}

extern "C" const char *sda_error_string(int code)
{
    switch (code) {
    case SDA_OK: return "ok";
    case SDA_E_INVALID: return "invalid argument";
    case SDA_E_BOUNDS: return "Bounds are note consistent min < max";
    case SDA_E_CUDA: return "CUDA error (see sda_last_cuda_error)";
    case SDA_E_WORKSPACE: return "workspace too small";
    case SDA_E_UNSUPPORTED: return "unsupported configuration";
    default: return "unknown error";
    }
}

extern "C" const char *sda_last_cuda_error(void) { return g_cuda_err; }

extern "C" int sda_functor_id(const char *name)
{
    if (!name) return SDA_E_INVALID;
    for (int i = 0; i < SDA_FN_COUNT; ++i)
        if (strcmp(name, kFunctorNames[i]) == 0) return i;
    return SDA_E_INVALID;
}

extern "C" const char *sda_functor_name(int functor)
{
    if (functor < 0 || functor >= SDA_FN_COUNT) return NULL;
    return kFunctorNames[functor];
}

extern "C" int sda_device_count(void)
{
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) return cuda_fail(e, "cudaGetDeviceCount");
    return n;
}

// ---- K1 on the host, the reference's formulas with libm                     :396-402, :76-86
static double host_temperature(double t0, double qv, long long i)
{
    const double t1 = exp((qv - 1) * log(2.0)) - 1.0;
    const double s = (double)i + 2.0;
    const double t2 = exp((qv - 1) * log(s)) - 1.0;
    return t0 * t1 / t2;
}
static double host_sigmax(double temperature, double qv)
{
    const double factor1 = exp(log(temperature) / (qv - 1.0));
    const double factor2 = exp((4.0 - qv) * log(qv - 1.0));
    const double factor3 = exp((2.0 - qv) * log(2.0) / (qv - 1.0));
    const double factor4 = sqrt(M_PI) * factor1 * factor2 / (factor3 * (3.0 - qv));
    const double factor5 = 1.0 / (qv - 1.0) - 0.5;
    const double d1 = 2.0 - factor5;
    const double factor6 = M_PI * (1.0 - factor5) / sin(M_PI * (1.0 - factor5)) / exp(lgamma(d1));
    return exp(-(qv - 1.0) * log(factor6 / factor4) / (3.0 - qv));
}

// number of K1 table entries the device needs: up to and including the first T < T_restart
static long long table_entries(const SdaConfig *c)
{
    if (c->temp_table && c->sigmax_table && c->table_len > 0) return c->table_len;
    long long n = 0;
    for (long long i = 0; i < c->maxiter; ++i) {
        ++n;
        if (host_temperature(c->initial_temp, c->visit, i) < c->temperature_restart) break;
    }
    return n > 0 ? n : 1;
}

// lanes per chain: the smallest power of two that leaves at most 8 coordinates per lane, so that
// D=50 runs 4 chains per warp with 7 coordinates per lane (89 % of the lanes busy in the O(D)
// loops instead of 78 % with one chain per warp) and warp-uniform work is shared by 32/G chains.
// Tape replay is serial by construction and keeps one chain per warp.
// Small batches (the suite runs 100 chains per job) cannot fill the machine at that packing, and
// the warp-wide local search serves the 32/G chains of a warp one after the other: G is widened
// while every chain group still gets resident lanes (148 SMs x 16 warps), up to one chain per warp.
static int pick_group(int dim, int rng_mode, long long n_chains)
{
    if (rng_mode == SDA_RNG_TAPE) return 32;
    const char *env = getenv("SDA_GROUP");
    if (env) {
        const int g = atoi(env);
        if (g == 1 || g == 2 || g == 4 || g == 8 || g == 16 || g == 32) return g;
    }
    int g = 1;
    while (g < 32 && (dim + g - 1) / g > 8) g <<= 1;
    while (g < 32 && n_chains * (2LL * g) <= 32LL * device_sms() * 16) g <<= 1;
    return g;
}

// a user functor is evaluated whole by every lane of its group, so fewer lanes per chain waste
// less: the smallest G whose 32/G chain rows (3*D doubles each) still fit 16 KB per warp
static int pick_group_user(int dim)
{
    int g = 1;
    while (g < 32 && (size_t)(32 / g) * 3 * dim * sizeof(double) > 16384) g <<= 1;
    return g;
}

struct LaunchPlan {
    int warps_per_block, blocks, group, sms;
    size_t smem;
    long long slots;
};

// rows_global: in/out.  In: -1 = decide here, 0/1 = forced.  Decision: the rows leave shared memory when
// a full block (8 warps) would not reach the target residency with them (D >~ 230 for one chain per warp).
static int plan_launch(int dim, long long n_chains, int pure_sa, int group, int min_blocks, LaunchPlan *lp,
                       int *rows_global = nullptr)
{
    int dev = 0, sms = 0, smem_max = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return cuda_fail(e, "cudaGetDevice");
    e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (e != cudaSuccess) return cuda_fail(e, "cudaDeviceGetAttribute");
    e = cudaDeviceGetAttribute(&smem_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    if (e != cudaSuccess) return cuda_fail(e, "cudaDeviceGetAttribute");
    int wpb = 8;
    const int ng = 32 / group;
    const size_t ls_per_warp = pure_sa ? 0 : (size_t)sda::ls_smem_doubles() * sizeof(double);
    // aim for SDA_ANNEAL_MIN_BLOCKS resident blocks: shrink the block before giving up residency
    const size_t budget = (size_t)smem_max / min_blocks - 1024;
    int rg = rows_global ? *rows_global : 0;
    if (rg < 0) {
        rg = (size_t)(4 + 3 * ng * wpb) * dim * sizeof(double) + wpb * ls_per_warp > budget ? 1 : 0;
        const char *env = getenv("SDA_ROWS_GLOBAL");
        if (env) rg = atoi(env) != 0;
    }
    if (rows_global) *rows_global = rg;
    if (rg && min_blocks == SDA_PHASED_MIN_BLOCKS) min_blocks = SDA_GROWS_MIN_BLOCKS;    // lean kernel, global rows
    const int rows_per_warp = rg ? 0 : 3 * ng;
    while (wpb > 1 && (size_t)(4 + rows_per_warp * wpb) * dim * sizeof(double) + wpb * ls_per_warp > budget) wpb >>= 1;
    size_t smem = (size_t)(4 + rows_per_warp * wpb) * dim * sizeof(double) + wpb * ls_per_warp;
    lp->group = group;
    lp->sms = sms;
    if (smem > (size_t)smem_max) return SDA_E_UNSUPPORTED;
    // resident blocks per SM: bounded by shared memory and by 2048 threads
    int bps = (int)((size_t)(smem_max) / (smem + 1024));
    if (bps < 1) bps = 1;
    int max_bps = 64 / wpb; // 64 warps per SM
    if (max_bps > min_blocks * 8 / wpb) max_bps = min_blocks * 8 / wpb; // register budget
    if (bps > max_bps) bps = max_bps;
    long long want = (n_chains + (long long)wpb * ng - 1) / ((long long)wpb * ng);
    long long blocks = (long long)sms * bps;
    if (blocks > want) blocks = want;
    if (blocks < 1) blocks = 1;
    lp->warps_per_block = wpb;
    lp->blocks = (int)blocks;
    lp->smem = smem;
    lp->slots = blocks * wpb;
    return SDA_OK;
}

static size_t align256(size_t v) { return (v + 255) & ~(size_t)255; }

// workspace layout: [counter 256B][temp table][sigmax table][xmin C*D][ls slots]
//                   [parked-chain buffers of the phased execution: recs, rec_x, ready_q, ls_q]
struct WsLayout {
    size_t temp, sig, xmin, ls, recs, recx, rq, lq, rows, total;
    long long ls_stride;
};

struct RunModeLite { int group; };
static RunModeLite run_group(const SdaProblem *p, const SdaConfig *c, long long n_chains);

static WsLayout workspace_layout(const SdaProblem *p, const SdaConfig *c, long long n_chains, long long slots)
{
    WsLayout w;
    memset(&w, 0, sizeof(w));
    const long long tn = table_entries(c);
    size_t off = 256;
    w.temp = off; off += align256((size_t)tn * sizeof(double));
    w.sig = off; off += align256((size_t)tn * sizeof(double));
    w.xmin = off; off += align256((size_t)n_chains * p->dim * sizeof(double));
    w.ls = off;
    w.ls_stride = (sda::ls_workspace_doubles(p->dim) + 31) & ~31LL;
    if (!c->pure_sa) {
        off += align256((size_t)slots * (size_t)w.ls_stride * sizeof(double));
        w.recs = off; off += align256((size_t)n_chains * sizeof(sda::ChainRec));
        w.recx = off; off += align256((size_t)n_chains * p->dim * sizeof(double));
        w.rq = off; off += align256((size_t)2 * n_chains * sizeof(int));
        w.lq = off; off += align256((size_t)n_chains * sizeof(int));
    }
    // chain rows of the large-D mapping (plan_launch decides whether they are used): every slot
    w.rows = off;
    {
        const RunModeLite g = run_group(p, c, n_chains);
        if ((size_t)(4 + 3 * (32 / g.group) * 8) * p->dim * sizeof(double) > 40000)     // may be chosen
            off += align256((size_t)slots * (size_t)(32 / g.group) * 3 * p->dim * sizeof(double));
    }
    w.total = off;
    return w;
}

// Phased execution (annealing launches alternating with local-search launches) is the default for
// large Philox batches with local search; SDA_PHASED=0 / 1 forces the persistent / phased path.
// Both paths run the same chain code and give bit-identical results.
#ifndef SDA_LS_MIN_BLOCKS
#define SDA_LS_MIN_BLOCKS 4
#endif
static bool use_phased(const SdaConfig *c, long long n_chains)
{
    if (c->pure_sa || c->rng_mode == SDA_RNG_TAPE || c->maxiter > 5000) return false;
    const char *env = getenv("SDA_PHASED");
    if (env) return atoi(env) == 1;
    // A round waits for its slowest local search, so the local-search launch needs many requests
    // per warp to average the heavy-tailed search lengths: measured crossover for D=50 at ~12 k
    // chains (8 k: -11 %, 16 k: +16 %, 32 k: +43 %, 64 k: +35 %); Sutton-Chen N=38 with 8 k chains
    // is 5x slower phased.  148 SMs x 4 CTAs x 4 warps = 2368 searches run concurrently.
    return n_chains >= 8LL * device_sms() * SDA_LS_MIN_BLOCKS * 4;
}

static RunModeLite run_group(const SdaProblem *p, const SdaConfig *c, long long n_chains)
{
    RunModeLite r;
    r.group = (p->functor == SDA_FN_USER && c->rng_mode != SDA_RNG_TAPE)
                  ? pick_group_user(p->dim) : pick_group(p->dim, c->rng_mode, n_chains);
    return r;
}

typedef void (*SdaKernelFn)(const DeviceArgs);

// A kernel's dynamic shared-memory ceiling is process-wide state of the function.  It is always set to
// the DEVICE's opt-in maximum, never to the size of the launch at hand: two host threads launching the
// same kernel for different problems raced on it ("too many resources requested for launch" in the
// thread whose larger request came first -- tests: test_concurrent_host_threads_with_different_problems).
// The ceiling does not affect occupancy; the size passed at launch does.
template <class F>
static cudaError_t allow_dynamic_smem(F fn, size_t need)
{
    int dev = 0, smem_max = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    e = cudaDeviceGetAttribute(&smem_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    if (e != cudaSuccess) return e;
    cudaFuncAttributes fa;
    e = cudaFuncGetAttributes(&fa, fn);
    if (e != cudaSuccess) return e;
    const int limit = smem_max - (int)fa.sharedSizeBytes;         // static + dynamic <= the opt-in maximum
    if (need > (size_t)limit) return cudaErrorInvalidValue;
    return cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, limit);
}

// SDA_LS_MEMO=0: every perturbed point of a gradient evaluated in full (A/B switch; same bits either way)
static int ls_memo_enabled(void)
{
    const char *env = getenv("SDA_LS_MEMO");
    return env ? (atoi(env) != 0) : 1;
}

static thread_local long long g_last_launches = 0;
extern "C" long long sda_last_launch_count(void) { return g_last_launches; }

// optional per-kernel timing (include/sdaopt_b200.h: sda_kernel_timing / sda_kernel_times)
struct TimedLaunch { cudaEvent_t a, b; int kind; };
static thread_local bool g_timing = false;
static thread_local std::vector<TimedLaunch> g_timed;
struct LaunchTimer {
    cudaStream_t s; int kind; bool on; TimedLaunch t;
    LaunchTimer(cudaStream_t s_, int kind_) : s(s_), kind(kind_), on(g_timing)
    {
        if (!on) return;
        t.kind = kind;
        if (cudaEventCreate(&t.a) != cudaSuccess || cudaEventCreate(&t.b) != cudaSuccess) { on = false; return; }
        cudaEventRecord(t.a, s);
    }
    ~LaunchTimer()
    {
        if (!on) return;
        cudaEventRecord(t.b, s);
        g_timed.push_back(t);
    }
};
extern "C" int sda_kernel_timing(int enable) { g_timing = enable != 0; return SDA_OK; }
extern "C" int sda_kernel_times(double ms[3], long long launches[3])
{
    if (!ms || !launches) return SDA_E_INVALID;
    for (int i = 0; i < 3; ++i) { ms[i] = 0.0; launches[i] = 0; }
    int rc = SDA_OK;
    for (TimedLaunch &t : g_timed) {
        float dt = 0.f;
        cudaError_t e = cudaEventSynchronize(t.b);
        if (e == cudaSuccess) e = cudaEventElapsedTime(&dt, t.a, t.b);
        if (e != cudaSuccess) rc = cuda_fail(e, "sda_kernel_times");
        else { ms[t.kind] += dt; launches[t.kind] += 1; }
        cudaEventDestroy(t.a);
        cudaEventDestroy(t.b);
    }
    g_timed.clear();
    return rc;
}

static int max_slots(void)
{
    // upper bound used for workspace sizing without touching the device: 160 SMs x 32 warps
    return 160 * 32;
}

extern "C" size_t sda_workspace_bytes(const SdaProblem *p, const SdaConfig *c, int64_t n_chains)
{
    if (!p || !c || p->dim <= 0 || n_chains <= 0) return 0;
    return workspace_layout(p, c, n_chains, max_slots()).total;
}

static int validate(const SdaProblem *p, const SdaConfig *c, int64_t n_chains)
{
    if (!p || !c || p->dim <= 0 || n_chains <= 0) return SDA_E_INVALID;
    if (!problem_shape_ok(p->functor, p->dim)) return SDA_E_INVALID;
    if (c->rng_mode != SDA_RNG_PHILOX && c->rng_mode != SDA_RNG_TAPE) return SDA_E_INVALID;
    if ((p->functor == SDA_FN_SHIFTED_RASTRIGIN || p->functor == SDA_FN_CONSTANT) &&
        (p->params == NULL || p->n_params < 1))
        return SDA_E_INVALID;
    return SDA_OK;
}

extern "C" int sda_batch_run(const SdaProblem *p, const SdaConfig *c, int64_t n_chains,
                             const double *x0, const uint64_t *seeds, const double *tape,
                             int64_t tape_len, const SdaResult *out, void *workspace,
                             size_t workspace_bytes, void *stream_)
{
    int rc = validate(p, c, n_chains);
    if (rc) return rc;
    if (!out || !out->x || !out->fun || !out->nit || !out->ncall || !out->status || !workspace)
        return SDA_E_INVALID;
    if (c->rng_mode == SDA_RNG_TAPE && (!tape || tape_len <= 0)) return SDA_E_INVALID;
    cudaStream_t stream = (cudaStream_t)stream_;
    LaunchPlan lp;
    const bool phased = use_phased(c, n_chains);
    const int group = run_group(p, c, n_chains).group;
    int rows_global = c->rng_mode == SDA_RNG_TAPE ? 0 : -1;      // tape replay is serial: rows stay in shared memory
    // pure SA on the production RNG: the lean instantiation (64 registers, 4 CTAs per SM) when several
    // chains share a warp or when the rows are in global memory anyway (large D: occupancy is what
    // hides their latency); one chain per warp with its rows in shared memory (128 < D <~ 230) keeps
    // the 128-register instantiation (Sutton-Chen N=128 measured +28 % in round 1)
    bool lean_sa = c->pure_sa && c->rng_mode != SDA_RNG_TAPE;
    rc = plan_launch(p->dim, n_chains, c->pure_sa || phased, group,
                     (phased || lean_sa) ? SDA_PHASED_MIN_BLOCKS : SDA_ANNEAL_MIN_BLOCKS, &lp, &rows_global);
    if (rc) return rc;
    if (lean_sa && group == 32 && !rows_global) {
        lean_sa = false;
        rows_global = -1;
        rc = plan_launch(p->dim, n_chains, 1, group, SDA_ANNEAL_MIN_BLOCKS, &lp, &rows_global);
        if (rc) return rc;
    }
    const WsLayout wl = workspace_layout(p, c, n_chains, max_slots());
    const size_t off_temp = wl.temp, off_sig = wl.sig, off_xmin = wl.xmin, off_ls = wl.ls;
    const size_t off_recs = wl.recs, off_recx = wl.recx, off_rq = wl.rq, off_lq = wl.lq;
    const long long ls_stride = wl.ls_stride;
    if (wl.total > workspace_bytes) return SDA_E_WORKSPACE;
    if (rows_global && wl.total == wl.rows) return SDA_E_WORKSPACE;      // layout did not reserve the rows
    char *ws = (char *)workspace;

    // K1 tables (host) -> device
    const long long tn = table_entries(c);
    std::vector<double> tt((size_t)tn), st((size_t)tn);
    if (c->temp_table && c->sigmax_table && c->table_len > 0) {
        memcpy(tt.data(), c->temp_table, (size_t)tn * sizeof(double));
        memcpy(st.data(), c->sigmax_table, (size_t)tn * sizeof(double));
    } else {
        for (long long i = 0; i < tn; ++i) {
            tt[(size_t)i] = host_temperature(c->initial_temp, c->visit, i);
            st[(size_t)i] = host_sigmax(tt[(size_t)i], c->visit);
        }
    }
    CK(cudaMemcpyAsync(ws + off_temp, tt.data(), (size_t)tn * sizeof(double), cudaMemcpyHostToDevice, stream));
    CK(cudaMemcpyAsync(ws + off_sig, st.data(), (size_t)tn * sizeof(double), cudaMemcpyHostToDevice, stream));
    CK(cudaMemsetAsync(ws, 0, 256, stream));
    // tt/st are pageable: cudaMemcpyAsync has staged them before it returned, so they may die here

    DeviceArgs a;
    memset(&a, 0, sizeof(a));
    a.functor = p->functor; a.dim = p->dim;
    a.lower = p->lower; a.upper = p->upper; a.params = p->params;
    a.maxiter = c->maxiter; a.initial_temp = c->initial_temp; a.qv = c->visit; a.qa = c->accept;
    a.maxfun = c->maxfun; a.temperature_restart = c->temperature_restart; a.pure_sa = c->pure_sa;
    a.max_calls = c->max_calls; a.fglob_tol = c->fglob + c->tol; a.trace_len = c->trace_len;
    a.temp_table = (const double *)(ws + off_temp);
    a.sigmax_table = (const double *)(ws + off_sig);
    a.table_len = tn;
    a.n_chains = n_chains; a.x0 = x0; a.seeds = (const unsigned long long *)seeds;
    a.tape = tape; a.tape_len = tape_len;
    a.out = *out;
    if (c->trace_len <= 0) { a.out.trace_e = NULL; a.out.trace_d = NULL; }
    a.xmin = (double *)(ws + off_xmin);
    a.ls_work = c->pure_sa ? NULL : (double *)(ws + off_ls);
    a.ls_stride = ls_stride;
    a.work_counter = (unsigned long long *)ws;
    int lsmax = p->dim * 6;
    if (lsmax < 100) lsmax = 100;
    if (lsmax > 1000) lsmax = 1000;
    a.ls_maxiter = lsmax;
    a.ls_memo = ls_memo_enabled();
    a.group = lp.group;
    a.rows = rows_global ? (double *)(ws + wl.rows) : NULL;

    const int threads = lp.warps_per_block * 32;
    if (phased) {
        // rounds of { annealing launch ; local-search launch }: a round advances every live chain to
        // its next local-search request, so maxiter rounds always suffice; launches that find their
        // queue empty return at once
        a.phased = 1;
        {
            const char *env = getenv("SDA_STEPS_PER_ROUND");
            a.steps_per_round = env ? atoi(env) : 8;
            if (a.steps_per_round < 1) a.steps_per_round = 1;
        }
        long long n_rounds = c->maxiter / a.steps_per_round + 48;
        {
            const char *env = getenv("SDA_ROUNDS");
            if (env) n_rounds = atoll(env);
        }
        a.recs = (sda::ChainRec *)(ws + off_recs);
        a.rec_x = (double *)(ws + off_recx);
        a.ready_q = (int *)(ws + off_rq);
        a.ls_q = (int *)(ws + off_lq);
        a.qctl = (long long *)(ws + 64);
        CK(cudaMemsetAsync(ws + 64, 0, 64, stream));
        const char *ls_env = getenv("SDA_LS_SYNC");
        const bool ls_sync = ls_env ? atoi(ls_env) == 1 : true;    // block-synchronous launch: default
        const int ls_wpb = ls_sync ? SDA_LS_SYNC_WARPS : 4;
        const size_t ls_smem = ((size_t)2 * p->dim + (size_t)ls_wpb * ((rows_global ? 0 : p->dim) + sda::ls_smem_doubles())) * sizeof(double);
        int ls_blocks = ls_sync ? lp.sms * SDA_LS_SYNC_BLOCKS : lp.sms * SDA_LS_MIN_BLOCKS;
        if ((long long)ls_blocks * ls_wpb > max_slots()) ls_blocks = max_slots() / ls_wpb;
        const SdaKernelFn k_anneal = rows_global ? sda::sda_anneal_kernel<false, sda::kPhasedAnneal, true>
                                                 : sda::sda_anneal_kernel<false, sda::kPhasedAnneal, false>;
        const SdaKernelFn k_ls = ls_sync ? (rows_global ? sda::sda_ls_phase_kernel<true, true> : sda::sda_ls_phase_kernel<true, false>)
                                         : (rows_global ? sda::sda_ls_phase_kernel<false, true> : sda::sda_ls_phase_kernel<false, false>);
        CK(allow_dynamic_smem(k_anneal, lp.smem));
        CK(allow_dynamic_smem(k_ls, ls_smem));
        const bool debug_rounds = getenv("SDA_DEBUG_ROUNDS") != nullptr;
        for (long long r = 0; r < n_rounds; ++r) {
            a.round = (int)r;
            {
                LaunchTimer tm(stream, 0);
                k_anneal<<<lp.blocks, threads, lp.smem, stream>>>(a);
            }
            sda::sda_rotate_kernel<<<1, 1, 0, stream>>>(a.qctl, 0);
            {
                LaunchTimer tm(stream, 1);
                k_ls<<<ls_blocks, ls_wpb * 32, ls_smem, stream>>>(a);
            }
            sda::sda_rotate_kernel<<<1, 1, 0, stream>>>(a.qctl, 1);
            if (debug_rounds) {     // SDA_DEBUG_ROUNDS=1: queue lengths per round on stderr (synchronises!)
                long long q[8];
                cudaStreamSynchronize(stream);
                cudaMemcpy(q, a.qctl, sizeof(q), cudaMemcpyDeviceToHost);
                if (r % 8 == 0 || r + 1 == n_rounds)
                    fprintf(stderr, "round %lld: ready for next %lld\n", r, q[sda::kReadyCount]);
            }
        }
        // whatever is still parked (chains that posted a request in almost every step) is finished
        // by the persistent kernel, resuming from the ready queue
        LaunchPlan lp2;
        int rg2 = rows_global;
        rc = plan_launch(p->dim, n_chains, 0, lp.group, SDA_ANNEAL_MIN_BLOCKS, &lp2, &rg2);
        if (rc) return rc;
        a.round = (int)n_rounds;
        const SdaKernelFn k_resume = rows_global ? sda::sda_anneal_kernel<false, sda::kPersistent, true>
                                                 : sda::sda_anneal_kernel<false, sda::kPersistent, false>;
        CK(allow_dynamic_smem(k_resume, lp2.smem));
        {
            LaunchTimer tm(stream, 2);
            k_resume<<<lp2.blocks, lp2.warps_per_block * 32, lp2.smem, stream>>>(a);
        }
        g_last_launches = 4 * n_rounds + 1;
    } else {
        SdaKernelFn k;
        if (c->rng_mode == SDA_RNG_TAPE) k = sda::sda_anneal_kernel<true, sda::kPersistent, false>;    // (rows never global)
        else if (lean_sa) k = rows_global ? sda::sda_anneal_kernel<false, sda::kPureSa, true>
                                          : sda::sda_anneal_kernel<false, sda::kPureSa, false>;
        else k = rows_global ? sda::sda_anneal_kernel<false, sda::kPersistent, true>
                             : sda::sda_anneal_kernel<false, sda::kPersistent, false>;
        CK(allow_dynamic_smem(k, lp.smem));
        LaunchTimer tm(stream, 0);
        k<<<lp.blocks, threads, lp.smem, stream>>>(a);
    }
    if (!phased) g_last_launches = 1;
    CK(cudaGetLastError());
    return SDA_OK;
}

// ---- small RAII helper for the host-pointer entry points -----------------------------------
struct DevBuf {
    void *p = nullptr;
    ~DevBuf() { if (p) cudaFree(p); }
    cudaError_t alloc(size_t bytes) { return cudaMalloc(&p, bytes ? bytes : 8); }
    cudaError_t upload(const void *h, size_t bytes)
    {
        cudaError_t e = alloc(bytes);
        if (e != cudaSuccess) return e;
        return cudaMemcpy(p, h, bytes, cudaMemcpyHostToDevice);
    }
};

static int upload_problem(const SdaProblem *p, SdaProblem *dp, DevBuf &lo, DevBuf &up, DevBuf &pa)
{
    for (int k = 0; k < p->dim; ++k)
        if (!(p->lower[k] < p->upper[k])) return SDA_E_BOUNDS;          // _sda.py:366-367
    CK(lo.upload(p->lower, sizeof(double) * p->dim));
    CK(up.upload(p->upper, sizeof(double) * p->dim));
    *dp = *p;
    dp->lower = (const double *)lo.p;
    dp->upper = (const double *)up.p;
    dp->params = NULL;
    if (p->params && p->n_params > 0) {
        CK(pa.upload(p->params, sizeof(double) * p->n_params));
        dp->params = (const double *)pa.p;
    }
    return SDA_OK;
}

extern "C" int sda_batch_run_host(const SdaProblem *p, const SdaConfig *c, int64_t n_chains,
                                  const double *x0, const uint64_t *seeds, const double *tape,
                                  int64_t tape_len, const SdaResult *out, int device)
{
    int rc = validate(p, c, n_chains);
    if (rc) return rc;
    if (!p->lower || !p->upper || !out) return SDA_E_INVALID;
    CK(cudaSetDevice(device));
    const size_t C = (size_t)n_chains, D = (size_t)p->dim;
    DevBuf lo, up, pa, dx0, dseeds, dtape, ws;
    SdaProblem dp;
    rc = upload_problem(p, &dp, lo, up, pa);
    if (rc) return rc;
    if (x0) CK(dx0.upload(x0, C * D * sizeof(double)));
    if (seeds) CK(dseeds.upload(seeds, C * sizeof(uint64_t)));
    if (tape && c->rng_mode == SDA_RNG_TAPE) CK(dtape.upload(tape, C * (size_t)tape_len * sizeof(double)));
    const size_t tl = c->trace_len > 0 ? (size_t)c->trace_len : 0;
    // one device block for all outputs
    DevBuf ox, ofun, onit, oncall, oncls, onls, ostat, ohit, onhit, ofhit, ote, otd, otu;
    CK(ox.alloc(C * D * 8)); CK(ofun.alloc(C * 8)); CK(onit.alloc(C * 8)); CK(oncall.alloc(C * 8));
    CK(oncls.alloc(C * 8)); CK(onls.alloc(C * 8)); CK(ostat.alloc(C * 4)); CK(ohit.alloc(C * 4));
    CK(onhit.alloc(C * 8)); CK(ofhit.alloc(C * 8)); CK(otu.alloc(C * 8));
    if (tl) { CK(ote.alloc(C * tl * 8)); CK(otd.alloc(C * tl)); CK(cudaMemset(ote.p, 0, C * tl * 8)); CK(cudaMemset(otd.p, 0, C * tl)); }
    SdaResult dout;
    dout.x = (double *)ox.p; dout.fun = (double *)ofun.p; dout.nit = (int64_t *)onit.p;
    dout.ncall = (int64_t *)oncall.p; dout.ncall_ls = (int64_t *)oncls.p; dout.n_ls = (int64_t *)onls.p;
    dout.status = (int32_t *)ostat.p; dout.hit = (int32_t *)ohit.p; dout.ncall_hit = (int64_t *)onhit.p;
    dout.f_hit = (double *)ofhit.p; dout.trace_e = tl ? (double *)ote.p : NULL;
    dout.trace_d = tl ? (uint8_t *)otd.p : NULL; dout.tape_used = (int64_t *)otu.p;
    const size_t wbytes = sda_workspace_bytes(&dp, c, n_chains);
    CK(ws.alloc(wbytes));
    rc = sda_batch_run(&dp, c, n_chains, (const double *)dx0.p, (const uint64_t *)dseeds.p,
                       (const double *)dtape.p, tape_len, &dout, ws.p, wbytes, NULL);
    if (rc) return rc;
    CK(cudaDeviceSynchronize());
#define DL(field, dev, bytes) \
    if (out->field) CK(cudaMemcpy(out->field, dev.p, bytes, cudaMemcpyDeviceToHost))
    DL(x, ox, C * D * 8); DL(fun, ofun, C * 8); DL(nit, onit, C * 8); DL(ncall, oncall, C * 8);
    DL(ncall_ls, oncls, C * 8); DL(n_ls, onls, C * 8); DL(status, ostat, C * 4); DL(hit, ohit, C * 4);
    DL(ncall_hit, onhit, C * 8); DL(f_hit, ofhit, C * 8); DL(tape_used, otu, C * 8);
    if (tl) { DL(trace_e, ote, C * tl * 8); DL(trace_d, otd, C * tl); }
#undef DL
    return SDA_OK;
}

extern "C" int sda_eval(const SdaProblem *p, const double *x, int64_t n, double *f, void *stream)
{
    if (!p || !x || !f || n <= 0 || p->dim <= 0) return SDA_E_INVALID;
    if (!problem_shape_ok(p->functor, p->dim)) return SDA_E_INVALID;
    long long blocks = (n + 7) / 8;
    if (blocks > device_sms() * 8) blocks = device_sms() * 8;
    sda::sda_eval_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(p->functor, p->dim, p->params, x, n, f,
                                                                       p->functor == SDA_FN_USER ? pick_group_user(p->dim) : pick_group(p->dim, SDA_RNG_PHILOX, 1LL << 40));
    CK(cudaGetLastError());
    return SDA_OK;
}

extern "C" int sda_eval_host(const SdaProblem *p, const double *x, int64_t n, double *f, int device)
{
    if (!p || !x || !f || n <= 0 || p->dim <= 0) return SDA_E_INVALID;
    if (!problem_shape_ok(p->functor, p->dim)) return SDA_E_INVALID;
    CK(cudaSetDevice(device));
    DevBuf dx, df, pa;
    SdaProblem dp = *p;
    if (p->params && p->n_params > 0) { CK(pa.upload(p->params, sizeof(double) * p->n_params)); dp.params = (const double *)pa.p; }
    CK(dx.upload(x, (size_t)n * p->dim * sizeof(double)));
    CK(df.alloc((size_t)n * sizeof(double)));
    int rc = sda_eval(&dp, (const double *)dx.p, n, (double *)df.p, NULL);
    if (rc) return rc;
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(f, df.p, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost));
    return SDA_OK;
}

extern "C" int sda_visiting_host(const SdaProblem *p, double qv, const double *x, int64_t step,
                                 double temperature, double sigmax, const double *tape,
                                 int64_t tape_len, double *x_visit, int64_t *tape_used, int device)
{
    (void)temperature;
    if (!p || !x || !tape || !x_visit || p->dim <= 0 || tape_len <= 0) return SDA_E_INVALID;
    CK(cudaSetDevice(device));
    DevBuf lo, up, pa, dx, dt, dv, du;
    SdaProblem dp;
    int rc = upload_problem(p, &dp, lo, up, pa);
    if (rc) return rc;
    CK(dx.upload(x, sizeof(double) * p->dim));
    CK(dt.upload(tape, sizeof(double) * (size_t)tape_len));
    CK(dv.alloc(sizeof(double) * p->dim));
    CK(du.alloc(8));
    DeviceArgs a;
    memset(&a, 0, sizeof(a));
    a.dim = p->dim; a.lower = dp.lower; a.upper = dp.upper; a.qv = qv;
    a.tape = (const double *)dt.p; a.tape_len = tape_len;
    sda::sda_visiting_kernel<<<1, 32, sizeof(double) * p->dim>>>(a, (const double *)dx.p, step, sigmax,
                                                                  (double *)dv.p, (long long *)du.p);
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(x_visit, dv.p, sizeof(double) * p->dim, cudaMemcpyDeviceToHost));
    if (tape_used) CK(cudaMemcpy(tape_used, du.p, 8, cudaMemcpyDeviceToHost));
    return SDA_OK;
}

extern "C" int sda_visit_fn_host(double qv, double sigmax, const double *tape, int64_t tape_len,
                                 double *out, int64_t n, int64_t *tape_used, int device)
{
    if (!tape || !out || n <= 0 || tape_len <= 0) return SDA_E_INVALID;
    CK(cudaSetDevice(device));
    DevBuf dt, dout, du;
    CK(dt.upload(tape, sizeof(double) * (size_t)tape_len));
    CK(dout.alloc(sizeof(double) * (size_t)n));
    CK(du.alloc(8));
    sda::sda_visit_fn_kernel<<<1, 32>>>(qv, sigmax, (const double *)dt.p, tape_len, (double *)dout.p, n,
                                        (long long *)du.p);
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(out, dout.p, sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost));
    if (tape_used) CK(cudaMemcpy(tape_used, du.p, 8, cudaMemcpyDeviceToHost));
    return SDA_OK;
}

extern "C" int sda_local_search_host(const SdaProblem *p, const double *x_in, int64_t n,
                                     double *x_out, double *f_out, int32_t *success, int64_t *nfev,
                                     int device)
{
    if (!p || !x_in || !x_out || !f_out || n <= 0 || p->dim <= 0) return SDA_E_INVALID;
    if (!problem_shape_ok(p->functor, p->dim)) return SDA_E_INVALID;
    CK(cudaSetDevice(device));
    DevBuf lo, up, pa, dxi, dxo, dfo, dsu, dnf, dls;
    SdaProblem dp;
    int rc = upload_problem(p, &dp, lo, up, pa);
    if (rc) return rc;
    const size_t D = (size_t)p->dim;
    CK(dxi.upload(x_in, (size_t)n * D * 8));
    CK(dxo.alloc((size_t)n * D * 8)); CK(dfo.alloc((size_t)n * 8)); CK(dsu.alloc((size_t)n * 4));
    CK(dnf.alloc((size_t)n * 8));
    int wpb = 4;
    int smem_max = 0;
    CK(cudaDeviceGetAttribute(&smem_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, device));
    const size_t lsb = (size_t)sda::ls_smem_doubles() * 8;
    while (wpb > 1 && (size_t)(2 + wpb) * D * 8 + wpb * lsb > (size_t)smem_max) wpb >>= 1;
    const size_t smem = (size_t)(2 + wpb) * D * 8 + wpb * lsb;
    long long blocks = (n + wpb - 1) / wpb;
    if (blocks > device_sms() * 4) blocks = device_sms() * 4;
    const long long stride = (sda::ls_workspace_doubles(p->dim) + 31) & ~31LL;
    CK(dls.alloc((size_t)blocks * wpb * (size_t)stride * 8));
    DeviceArgs a;
    memset(&a, 0, sizeof(a));
    a.functor = p->functor; a.dim = p->dim; a.lower = dp.lower; a.upper = dp.upper; a.params = dp.params;
    a.ls_work = (double *)dls.p; a.ls_stride = stride;
    int lsmax = p->dim * 6;
    if (lsmax < 100) lsmax = 100;
    if (lsmax > 1000) lsmax = 1000;
    a.ls_maxiter = lsmax;
    a.ls_memo = ls_memo_enabled();
    CK(allow_dynamic_smem(sda::sda_local_search_kernel, smem));
    sda::sda_local_search_kernel<<<(int)blocks, wpb * 32, smem>>>(a, (const double *)dxi.p, n, (double *)dxo.p,
                                                                   (double *)dfo.p, (int *)dsu.p, (long long *)dnf.p);
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(x_out, dxo.p, (size_t)n * D * 8, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(f_out, dfo.p, (size_t)n * 8, cudaMemcpyDeviceToHost));
    if (success) CK(cudaMemcpy(success, dsu.p, (size_t)n * 4, cudaMemcpyDeviceToHost));
    if (nfev) CK(cudaMemcpy(nfev, dnf.p, (size_t)n * 8, cudaMemcpyDeviceToHost));
    return SDA_OK;
}

extern "C" int sda_energy_reset_host(const SdaProblem *p, const double *x0, const double *tape,
                                     int64_t tape_len, int32_t have_best, double *current_location,
                                     double *current_energy, double *xbest, double *ebest,
                                     int32_t *status, int64_t *tape_used, int device)
{
    if (!p || !tape || tape_len <= 0 || !current_location || !current_energy || !xbest || !ebest || !status)
        return SDA_E_INVALID;
    if (!problem_shape_ok(p->functor, p->dim)) return SDA_E_INVALID;
    if ((p->functor == SDA_FN_SHIFTED_RASTRIGIN || p->functor == SDA_FN_CONSTANT) && (p->params == NULL || p->n_params < 1))
        return SDA_E_INVALID;
    CK(cudaSetDevice(device));
    DevBuf lo, up, pa, dx0, dt, dxc, dxb, dsc, dst, dtu;
    SdaProblem dp;
    int rc = upload_problem(p, &dp, lo, up, pa);
    if (rc) return rc;
    const size_t D = (size_t)p->dim;
    if (x0) CK(dx0.upload(x0, D * 8));
    CK(dt.upload(tape, sizeof(double) * (size_t)tape_len));
    CK(dxc.alloc(D * 8));
    CK(dxb.upload(xbest, D * 8));
    const double sc[2] = { 0.0, *ebest };
    CK(dsc.upload(sc, sizeof(sc)));
    CK(dst.alloc(4)); CK(dtu.alloc(8));
    DeviceArgs a;
    memset(&a, 0, sizeof(a));
    a.functor = p->functor; a.dim = p->dim; a.lower = dp.lower; a.upper = dp.upper; a.params = dp.params;
    a.tape = (const double *)dt.p; a.tape_len = tape_len;
    a.fglob_tol = nan("");
    const size_t smem = 5 * D * sizeof(double);
    CK(allow_dynamic_smem(sda::sda_energy_reset_kernel, smem));
    sda::sda_energy_reset_kernel<<<1, 32, smem>>>(a, (const double *)dx0.p, have_best ? 1 : 0, (double *)dxc.p,
                                                  (double *)dxb.p, (double *)dsc.p, (int *)dst.p, (long long *)dtu.p);
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    double sc_out[2];
    CK(cudaMemcpy(sc_out, dsc.p, sizeof(sc_out), cudaMemcpyDeviceToHost));
    *current_energy = sc_out[0];
    *ebest = sc_out[1];
    CK(cudaMemcpy(current_location, dxc.p, D * 8, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(xbest, dxb.p, D * 8, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(status, dst.p, 4, cudaMemcpyDeviceToHost));
    if (tape_used) CK(cudaMemcpy(tape_used, dtu.p, 8, cudaMemcpyDeviceToHost));
    return SDA_OK;
}

extern "C" int sda_math_probe_host(int kind, const double *x, int64_t n, double *out, int device)
{
    if (!x || !out || n <= 0 || kind < 0 || kind > 4) return SDA_E_INVALID;
    CK(cudaSetDevice(device));
    DevBuf dx, dy;
    CK(dx.upload(x, sizeof(double) * n));
    CK(dy.alloc(sizeof(double) * n));
    sda::sda_math_probe_kernel<<<(unsigned)((n + 255) / 256), 256>>>(kind, (const double *)dx.p, n, (double *)dy.p);
    CK(cudaGetLastError());
    CK(cudaMemcpy(out, dy.p, sizeof(double) * n, cudaMemcpyDeviceToHost));
    return SDA_OK;
}

extern "C" int sda_fp64_peak(double *tflops, double *ms_out, int device)
{
    if (!tflops) return SDA_E_INVALID;
    CK(cudaSetDevice(device));
    int sms = 0;
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
    const int blocks = sms * 8, threads = 256, iters = 1 << 16;
    DevBuf out;
    CK(out.alloc((size_t)blocks * threads * 8));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    sda::sda_dfma_kernel<<<blocks, threads>>>((double *)out.p, 1 << 12);   // warm-up
    float best = 1e30f;
    for (int rep = 0; rep < 3; ++rep) {
        CK(cudaEventRecord(e0));
        sda::sda_dfma_kernel<<<blocks, threads>>>((double *)out.p, iters);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms = 0.f;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) best = ms;
    }
    CK(cudaGetLastError());
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    const double flops = 2.0 * 8.0 * (double)iters * (double)blocks * threads;
    *tflops = flops / (best * 1e-3) / 1e12;
    if (ms_out) *ms_out = best;
    return SDA_OK;
}
