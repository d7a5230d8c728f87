// jtb_scout.cuh — depth-first scouts that run beside the breadth-first crowd of jtb_wgl.cuh.
//
// Why: the persistent search kernel expands everything reachable, thousands of configs at a time.  On a VALID
// history with many crashed (:info) ops the reachable space is astronomically large (every crashed write may be
// linearized almost anywhere) and the crowd runs out of budget, while knossos.wgl's single depth-first thread walks
// straight to a linearization because it always tries the earliest-invoked entry first (soak findings
// in DESIGN.md §7).  A scout is ONE warp that performs exactly that walk on the device: same config keys, same
// model step, same candidate rules as the crowd, but children are tried one at a time in a fixed priority order,
// with a private visited table (the crowd's table would block the scout's path with configs that are merely
// queued) and an explicit stack in HBM.  Order 0 is knossos' order; the other scouts try different orders of the
// same candidates (a portfolio: depth-first luck depends on the order).
//
// A scout can only ever report VALID (it found a complete linearization).  It never reports INVALID: exhausting
// one order's budget proves nothing, and the crowd's exhaustive count stays the quantity the parity tests compare.
// Scouts claim (shard, order) pairs from a counter, so a multi-key history is swept shard by shard.
#pragma once
#include "jtb_wgl.cuh"

namespace jtb {

static_assert(sizeof(ClassRec) == 32, "class records are read as two int4");
constexpr int SCOUT_ORDERS = 4;
constexpr int SCOUT_SMEM_CLASSES = 256;   // class records of the current shard kept in shared memory (8 KB)
constexpr int SCOUT_CTL_WORDS = 8;   // [0] next pair, [1] stop (host), [2] steps, [3] inserts, [4] shards decided

struct ScoutParams {
    const uint64_t* init;      // n_init initial entries (EW words each), one per searchable shard
    int n_init;
    int n_orders;
    uint64_t* tables;          // one private visited table per scout: (slot_mask + 1) * KW words each
    uint64_t slot_mask;
    uint64_t* stacks;          // one stack per scout: stack_cap frames of EW + 1 words (entry, cursor)
    uint32_t stack_cap;
    unsigned long long* ctl;   // SCOUT_CTL_WORDS
    unsigned long long pair_budget;   // expansion steps a scout spends on one (shard, order) pair
};

// Priority of a candidate (smaller is tried first; unique within a config because positions are unique).
//   order 0: invocation order, crashed ops interleaved        (knossos.wgl's entry-list order)
//   order 1: completed ops in invocation order, then crashed ops
//   order 2: the op whose return is the frontier first, then as order 1
//   order 3: latest-invoked completed op first, then crashed ops (earliest first)
__device__ __forceinline__ uint32_t scout_prio(int order, uint32_t inv_pos, bool is_front, bool crashed) {
    const uint32_t late = crashed ? (1u << 30) : 0u;
    switch (order) {
    case 0: return inv_pos;
    case 1: return inv_pos + late;
    case 2: return is_front ? 0u : 1u + inv_pos + late;
    default: return crashed ? (1u << 30) + inv_pos : (1u << 30) - 1u - inv_pos;
    }
}

template <int MODEL, int KW, bool EAGER>
__global__ void __launch_bounds__(32) wgl_scout_kernel(const WglParams p, const ScoutParams sp, const int neg_ok) {
    constexpr int SW = MODEL == JTB_MODEL_BANK ? 12 : MODEL == JTB_MODEL_SET ? 8 : 4;
    constexpr bool BANK = MODEL == JTB_MODEL_BANK;
    constexpr int EW = KW + (BANK ? 4 : 0);
    constexpr int FW = EW + 1;
    constexpr unsigned FULL = 0xffffffffu;
    constexpr uint32_t NONE = 0xffffffffu;
    constexpr int CLASS_ID = 1 << 16;
    __shared__ int4 s_cls[2 * SCOUT_SMEM_CLASSES];   // (op, {first, n, word, shift|width<<8}) per class
    const int lane = threadIdx.x;
    uint64_t* const table = sp.tables + (size_t)blockIdx.x * (sp.slot_mask + 1) * KW;
    uint64_t* const stack = sp.stacks + (size_t)blockIdx.x * sp.stack_cap * FW;
    Ctrl* ctrl = p.ctrl;
    const int cand_rounds = p.S_pad / 32;
    const int cls_rounds = (p.max_nc + 31) / 32;
    const unsigned long long insert_cap = (sp.slot_mask + 1) / 2;
    const unsigned long long n_pairs = (unsigned long long)sp.n_init * (unsigned long long)sp.n_orders;
    unsigned long long inserts = 0, steps_total = 0, decided = 0;
    bool quit = false;

    while (!quit) {
        unsigned long long q = 0;
        if (lane == 0) q = atomicAdd(&sp.ctl[0], 1ull);
        q = __shfl_sync(FULL, q, 0);
        if (q >= n_pairs) break;
        const int order = (int)(q / (unsigned long long)sp.n_init);
        const uint64_t* ie = sp.init + (size_t)(q % (unsigned long long)sp.n_init) * EW;
        uint64_t w[KW];
        int32_t pbal[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
        for (int i = 0; i < KW; ++i) w[i] = __ldg(ie + i);
        if constexpr (BANK) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const uint64_t v = __ldg(ie + KW + i);
                pbal[2 * i] = (int32_t)(uint32_t)v;
                pbal[2 * i + 1] = (int32_t)(uint32_t)(v >> 32);
            }
        }
        if (lane == 0)
            for (int i = 0; i < EW; ++i) __stcg(stack + i, __ldg(ie + i));
        uint32_t cursor = 0;   // smallest priority not yet tried at the current config
        int top = 0;           // the current config is stack[top]
        unsigned long long steps = 0;
        unsigned visits = 0;
        // the shard's class records: one dependent global round trip less per step when they sit in shared memory
        const int4* cls_tab;
        {
            const int32_t* row0 = p.rows + (size_t)(int)((w[0] >> 32) & RANK_MASK) * p.row_words;
            const int cb = __ldg(row0 + 11), nc = __ldg(row0 + 12);
            cls_tab = reinterpret_cast<const int4*>(p.classes + cb);
            if (nc <= SCOUT_SMEM_CLASSES) {
                __syncwarp();
                for (int i = lane; i < 2 * nc; i += 32) s_cls[i] = __ldg(cls_tab + i);
                __syncwarp();
                cls_tab = s_cls;
            }
        }

        for (;;) {
            if ((visits++ & 15) == 0) {
                int s = 0;
                if (lane == 0) s = ld_volatile(&sp.ctl[1]) != 0 || ld_volatile(&ctrl->n_undecided) <= 0;
                if (__shfl_sync(FULL, s, 0)) { quit = true; break; }
            }
            if (inserts >= insert_cap) { quit = true; break; }
            if (++steps > sp.pair_budget) break;
            // ---- the configuration's frontier row ------------------------------------------------------
            const int gj = (int)((w[0] >> 32) & RANK_MASK);
            const int32_t preg = (int32_t)(uint32_t)w[0];
            const int32_t* row = p.rows + (size_t)gj * p.row_words;
            const int32_t extra = __ldg(row + (lane & 15));
            const int fr_pos = __shfl_sync(FULL, extra, 8);
            const int shard = __shfl_sync(FULL, extra, 9);
            const int gj_end = __shfl_sync(FULL, extra, 10);
            const int ncls = __shfl_sync(FULL, extra, 12);
            const int rslot = __shfl_sync(FULL, extra, 13);
            {
                int f = 0;
                if (lane == 0) f = ld_volatile(&p.shard_found[shard]);
                if (__shfl_sync(FULL, f, 0)) break;   // somebody decided this shard
            }
            // ---- evaluate every candidate ONCE per visit; lane-private priorities (NONE = not a consistent child) ----
            constexpr int SR = 2, CR = 4;     // cached rounds: 64 slots, 128 classes (more classes: re-evaluate)
            uint32_t prs[SR], prc[CR];
            uint32_t xbest = NONE, ebest = NONE;   // xbest: best of the un-cached class rounds (filtered by cursor)
            int xbest_id = -1, ebest_id = -1;
#pragma unroll
            for (int r = 0; r < SR; ++r) {
                prs[r] = NONE;
                if (r < cand_rounds) {
                    const int t = r * 32 + lane;
                    const int32_t* cell = row + ROW_EXTRA + t * SW;
                    const int4 op = __ldg(reinterpret_cast<const int4*>(cell));
                    const bool cand = op.x >= 0 && !((w[1] >> t) & 1ull);
                    int32_t creg = preg;
                    int32_t cbal[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) cbal[i] = pbal[i];
                    const bool ok = cand && model_step<MODEL>(op, creg, cbal, cell, neg_ok != 0, w[1]);
                    if (ok) {
                        const bool is_read = (op.x & 0xff) == JTB_F_READ;
                        const uint32_t ipos = (uint32_t)((BANK && !is_read) ? __ldg(cell + 4) : op.w);
                        if (EAGER && is_read && ipos < ebest) { ebest = ipos; ebest_id = t; }
                        prs[r] = scout_prio(order, ipos, t == rslot, false);
                    }
                }
            }
#pragma unroll
            for (int r = 0; r < CR; ++r) prc[r] = NONE;
            for (int r = 0; r < cls_rounds; ++r) {
                const int c = r * 32 + lane;
                bool cand = c < ncls;
                int first = 0, n = 0, word = 1, shift_width = 0;
                int4 cop = make_int4(OP_IMPOSSIBLE, 0, 0, 0);
                if (cand) {
                    const int4 b = cls_tab[2 * c + 1];
                    cop = cls_tab[2 * c];
                    first = b.x; n = b.y; word = b.z; shift_width = b.w;
                }
                const int shift = shift_width & 0xff, width = shift_width >> 8;
                uint64_t field = 0;
#pragma unroll
                for (int i = 0; i < KW; ++i) if (i == word) field = w[i];
                const int count = (int)((field >> shift) & ((1ull << width) - 1));
                cand = cand && count < n;
                uint32_t ipos = 0;
                if (cand) {
                    ipos = (uint32_t)__ldg(p.cls_inv_pos + first + count);
                    cand = (int)ipos < fr_pos;
                }
                int32_t creg = preg;
                int32_t cbal[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) cbal[i] = pbal[i];
                const bool ok = cand && model_step<MODEL>(cop, creg, cbal, nullptr, neg_ok != 0, w[1]);
                const uint32_t pr = ok ? scout_prio(order, ipos, false, true) : NONE;
                if (r < CR) {
#pragma unroll
                    for (int k = 0; k < CR; ++k) if (k == r) prc[k] = pr;
                } else if (pr >= cursor && pr < xbest) { xbest = pr; xbest_id = CLASS_ID + c; }
            }
            const bool cached = cls_rounds <= CR;   // all priorities are in registers: siblings need no re-evaluation
            const uint32_t emn = EAGER ? __reduce_min_sync(FULL, ebest) : NONE;
            bool leave_pair = false;
            // ---- siblings in priority order: the first child that is new becomes the current config ---------------
            for (;;) {
                int pick = -1;
                uint32_t next_cursor = NONE;
                if (emn != NONE) {
                    // eager reads: the earliest-invoked consistent read is the ONLY child of this config
                    if (cursor == 0) {
                        const unsigned who = __ballot_sync(FULL, ebest == emn);
                        pick = __shfl_sync(FULL, ebest_id, __ffs(who) - 1);
                    }
                } else {
                    uint32_t best = xbest;
                    int best_id = xbest_id;
#pragma unroll
                    for (int r = 0; r < SR; ++r)
                        if (prs[r] != NONE && prs[r] >= cursor && prs[r] < best) { best = prs[r]; best_id = r * 32 + lane; }
#pragma unroll
                    for (int r = 0; r < CR; ++r)
                        if (prc[r] != NONE && prc[r] >= cursor && prc[r] < best) { best = prc[r]; best_id = CLASS_ID + r * 32 + lane; }
                    const uint32_t mn = __reduce_min_sync(FULL, best);
                    if (mn != NONE) {
                        const unsigned who = __ballot_sync(FULL, best == mn);
                        pick = __shfl_sync(FULL, best_id, __ffs(who) - 1);
                        next_cursor = mn + 1;
                    }
                }
                if (pick < 0) {
                    // ---- every child tried: backtrack ---------------------------------------------------
                    if (top == 0) { leave_pair = true; break; }   // order exhausted without a linearization: no verdict
                    --top;
                    __syncwarp();
                    const uint64_t* fr = stack + (size_t)top * FW;
#pragma unroll
                    for (int i = 0; i < KW; ++i) w[i] = __ldcg(fr + i);
                    if constexpr (BANK) {
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const uint64_t v = __ldcg(fr + KW + i);
                            pbal[2 * i] = (int32_t)(uint32_t)v;
                            pbal[2 * i + 1] = (int32_t)(uint32_t)(v >> 32);
                        }
                    }
                    cursor = (uint32_t)__ldcg(fr + EW);
                    break;
                }
                // ---- build the child (all lanes compute the same values) ----------------------------------
                uint64_t cw[KW];
#pragma unroll
                for (int i = 0; i < KW; ++i) cw[i] = w[i];
                int32_t creg = preg;
                int32_t cbal[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) cbal[i] = pbal[i];
                int cgj = gj;
                if (pick < CLASS_ID) {
                    const int t = pick;
                    const int32_t* cell = row + ROW_EXTRA + t * SW;
                    const int4 op = __ldg(reinterpret_cast<const int4*>(cell));
                    model_step<MODEL>(op, creg, cbal, cell, neg_ok != 0, w[1]);
                    if (t == rslot) {
                        // the frontier op is linearized: the frontier passes every already-linearized return
                        uint64_t m = w[1];
                        int adv = 0;
                        const int32_t* rw = row;
                        int32_t ex = extra;
                        for (;;) {
                            const int32_t word = __shfl_sync(FULL, ex, lane >> 2);
                            const int sl = (word >> (8 * (lane & 3))) & 0xff;
                            const bool setb = sl != 0xff && ((m >> sl) & 1ull);
                            const unsigned peers = __match_any_sync(FULL, sl);
                            const bool pass = setb && (peers & ((1u << lane) - 1)) == 0;
                            const unsigned pm = __ballot_sync(FULL, pass);
                            const int n = pm == FULL ? 32 : __ffs(~pm) - 1;
                            const uint64_t clr = (lane < n) ? (1ull << sl) : 0ull;
                            const uint32_t clo = __reduce_or_sync(FULL, (uint32_t)clr);
                            const uint32_t chi = __reduce_or_sync(FULL, (uint32_t)(clr >> 32));
                            m &= ~((uint64_t)clo | ((uint64_t)chi << 32));
                            adv += n;
                            if (n < 32) break;
                            rw += (size_t)32 * p.row_words;
                            ex = __ldg(rw + (lane & 15));
                        }
                        cgj = gj + 1 + adv;
                        cw[1] = m;
                    } else {
                        cw[1] |= 1ull << t;
                    }
                } else {
                    const int4 b = cls_tab[2 * (pick - CLASS_ID) + 1];
                    const int4 cop = cls_tab[2 * (pick - CLASS_ID)];
                    model_step<MODEL>(cop, creg, cbal, nullptr, neg_ok != 0, w[1]);
                    const int shift = b.w & 0xff;
#pragma unroll
                    for (int i = 1; i < KW; ++i) if (i == b.z) cw[i] += 1ull << shift;
                }
                cw[0] = KEY_VALID | ((uint64_t)(uint32_t)cgj << 32) |
                        ((BANK || MODEL == JTB_MODEL_SET) ? 0ull : (uint64_t)(uint32_t)creg);
                cursor = next_cursor;
                if (cgj >= gj_end) {
                    // every :ok op of the shard is linearized -> VALID (same bookkeeping as the crowd)
                    if (lane == 0 && atomicExch(&p.shard_found[shard], 1) == 0) {
                        ++decided;
                        const int left = atomicSub(&ctrl->n_undecided, 1);
                        __threadfence();
                        if (left == 1) atomicCAS(&ctrl->stop, 0, 1);
                    }
                    leave_pair = true;
                    break;
                }
                int res = 0;
                if (lane == 0) {
                    int plen;
                    res = table_insert<KW>(table, sp.slot_mask, cw, &plen, /*cas_first=*/true);   // private table: one round trip
                }
                res = __shfl_sync(FULL, res, 0);
                if (res < 0) { quit = true; leave_pair = true; break; }
                if (res == 0) {
                    // seen before: next sibling (straight from the cached priorities when they are complete)
                    if (!cached) break;
                    ++steps;
                    continue;
                }
                ++inserts;
                if ((uint32_t)(top + 1) >= sp.stack_cap) { leave_pair = true; break; }   // cannot happen: depth <= ops
                if (lane == 0) {
                    __stcg(stack + (size_t)top * FW + EW, (uint64_t)cursor);   // where the parent resumes
                    uint64_t* fr = stack + (size_t)(top + 1) * FW;
#pragma unroll
                    for (int i = 0; i < KW; ++i) __stcg(fr + i, cw[i]);
                    if constexpr (BANK) {
#pragma unroll
                        for (int i = 0; i < 4; ++i)
                            __stcg(fr + KW + i, (uint64_t)(uint32_t)cbal[2 * i] | ((uint64_t)(uint32_t)cbal[2 * i + 1] << 32));
                    }
                }
                ++top;
#pragma unroll
                for (int i = 0; i < KW; ++i) w[i] = cw[i];
#pragma unroll
                for (int i = 0; i < 8; ++i) pbal[i] = cbal[i];
                cursor = 0;
                break;
            }
            if (leave_pair) break;
        }
        steps_total += steps;
    }
    if (lane == 0) {
        atomicAdd(&sp.ctl[2], steps_total);
        atomicAdd(&sp.ctl[3], inserts);
        atomicAdd(&sp.ctl[4], decided);
    }
}

// Re-arms a paused search: clears stop/cause unless the scouts decided every shard in the meantime.
// (scouts: n_undecided first, then stop; here: stop first, then n_undecided — one of the two always sees the other)
__global__ void wgl_resume_ctrl_kernel(Ctrl* ctrl) {
    ctrl->cause = 0;
    atomicExch(&ctrl->stop, 0);
    __threadfence();
    if (ld_volatile(&ctrl->n_undecided) <= 0) atomicCAS(&ctrl->stop, 0, 1);
}

}  // namespace jtb
