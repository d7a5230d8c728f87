// jtb_level.cuh — the LEVEL engine: the Wing–Gong/Lowe configuration search as a level-synchronous sweep.
//
// Replaces the hot loop of knossos.wgl/analysis (`step` over every call entry that may be linearized
// next, then `cache.add((linearized BitSet, model))`) for exhaustive searches.  Same configurations, same keys, same
// per-thread expansion core (jtb_expand.h) as the work-list engines (jtb_wgl.cuh / jtb_search.cuh); what differs is the
// ORDER, and what that order buys on a B200:
//
//   depth(config) = number of linearized ops = frontier rank + popcount(open-slot mask) + crashed-class counts is a
//   function of the key, and every move adds exactly one.  So two equal configurations always meet IN THE SAME LEVEL:
//   the visited set only has to live for one level.  The engine keeps two level arrays (ping-pong, coalesced) and ONE
//   small hash window sized to the level (16 slots per configuration, 6-bit epoch tag in every slot, stale entries are
//   simply overwritten: no clearing, no growth, no re-hash, no pause/resume).  At the bench sizes the window is a few
//   MB — the probe stream runs out of the 126 MB L2 (measured 287 G random 16 B probes/s, profiles/r2_probe_sweep.json)
//   instead of HBM (36.6 G/s) — and memory no longer bounds the search: 10^10-configuration spaces fit.
//
//   Inside a level, a warp takes 32 configurations: phase 1, one lane per configuration, finds the candidate ops
//   (bit masks from the frontier row); phase 2 hands EVERY child of the 32 configurations to its own lane (prefix sum
//   + k-th-set-bit select), so all probes of the warp are in flight at once and no lane idles on a configuration with
//   fewer children; new children are compacted in shared memory and appended to the next level 32 at a time
//   (one atomic per 32 entries, fully coalesced stores).
//
//   Levels are separated by one grid barrier (monotone counter in HBM).  Narrow levels (<= narrow_max
//   configurations: the start of every search, eager-read searches) are run by CTA 0 alone with __syncthreads()
//   between them while the other CTAs wait at the barrier.
//
// Verdict / witness / configuration count are those of the work-list engines: the first configuration whose frontier
// passes the shard's last return => VALID; an empty level => INVALID with witness = furthest frontier reached;
// configs = sum of the level sizes (every distinct configuration is inserted exactly once).
#pragma once
#include "jtb_expand.h"
#include "jtb_wgl.cuh"

namespace jtb {

#ifndef JTB_LV_WARPS
#define JTB_LV_WARPS 8
#endif
#ifndef JTB_LV_CTAS
#define JTB_LV_CTAS 3
#endif
constexpr int LV_WARPS = JTB_LV_WARPS;
constexpr int LV_THREADS = LV_WARPS * 32;
constexpr int LV_STAGE = 64;        // staged new entries per warp (ring; flushed 32 at a time)
constexpr int LV_MAX_PROBE = 128;
// table slots only: bits 56..61 of word 0 hold the epoch of the insertion (a key's rank must stay below 2^24)
constexpr uint64_t LV_TAG_MASK = 0x3full << 56;
constexpr int64_t LV_MAX_RANKS = 1ll << 24;

struct LvSlot {   // what one attempt (a level, or its repetition with a larger window) produced; three in rotation
    alignas(128) unsigned long long cnt;   // entries appended to the next level
    int stop;      // 1: every shard is decided   2: give up (cause)
    int cause;
    int retry;     // a probe sequence ran off the window: repeat the level with a larger one
    int pad;
};
struct LvState {   // identical in every thread of the grid
    unsigned long long level, attempt, n_in, total;
    int epoch, in_idx, boost, stop, cause;
    unsigned long long zeroed;   // table slots known to be initialised
};
struct LvCtrl {
    alignas(128) unsigned long long bar;     // grid barrier: arrivals, monotone
    alignas(128) LvSlot slot[3];
    alignas(128) LvState pub;                // state after a run of narrow levels (CTA 0 -> everyone)
    alignas(128) int n_undecided;
    alignas(128) unsigned long long probes;
    unsigned long long max_probe_len, max_width, max_window, narrow_levels, retries, t0, t1;
    LvState fin;
};

struct LvParams {
    const int32_t* rows;
    const ClassRec* classes;
    const int32_t* cls_inv_pos;
    uint64_t* table;
    uint64_t table_slots;     // capacity (power of two)
    uint64_t* buf[2];         // level arrays, entries of EW words
    uint64_t buf_cap;         // entries per array
    LvCtrl* ctrl;
    int* shard_found;
    int* shard_max_rank;
    int row_words, sum_off, n_shards;
    unsigned long long max_configs;
    unsigned long long time_budget_ns;
    LvState init;              // where to start (level 0, or the level a grown relaunch resumes at)
    uint32_t narrow_max;       // a level of at most this many configurations is run by CTA 0 alone
    uint32_t slots_per_config; // window = pow2ceil(n_in * slots_per_config), at least min_slots
    uint64_t min_slots;
};

template <int KW, int EW, bool BAL>
struct LvScratch {   // per warp, shared memory
    uint64_t w[32][KW];
    uint64_t todo[32], cls_todo[32], rd_ok[32];
    int32_t bal[BAL ? 32 : 1][8];
    int32_t hdr[32][6];     // fr_pos, shard, gj_end, cls_base, rslot, ncls
    uint32_t start[36];     // exclusive prefix of the child counts; [32] = total
    uint64_t stage[LV_STAGE][EW];
};

__device__ __forceinline__ int select64(uint64_t m, int k) {   // position of the k-th (0-based) set bit
    const uint32_t lo = (uint32_t)m;
    const int c = __popc(lo);
    return k < c ? (int)__fns(lo, 0, k + 1) : 32 + (int)__fns((uint32_t)(m >> 32), 0, k - c + 1);
}

__device__ __forceinline__ void lv_load_state(LvState& dst, const LvState* src) {   // through L2, word by word
    static_assert(sizeof(LvState) % 8 == 0, "LvState is copied as 64-bit words");
    uint64_t* d = reinterpret_cast<uint64_t*>(&dst);
    const uint64_t* q = reinterpret_cast<const uint64_t*>(src);
#pragma unroll
    for (int i = 0; i < (int)(sizeof(LvState) / 8); ++i) d[i] = ldcg64(q + i);
}

// Probe + insert into the level window.  1 = inserted (new this level), 0 = already present, -1 = window exhausted.
// A slot is free when it is empty or carries another epoch's tag (dead: every level only ever looks for its own keys).
template <int KW>
__device__ __forceinline__ int lv_insert(uint64_t* table, uint64_t mask, const uint64_t (&k)[KW], uint64_t tag, int* plen) {
    uint64_t idx = hash_key<KW>(k) & mask;
    const uint64_t mine0 = k[0] | tag;
    for (int i = 0; i < LV_MAX_PROBE; ++i) {
        uint64_t* slot = table + idx * KW;
        K128 cur = ldcg128(slot);
        for (int round = 0; round < 2; ++round) {
            const bool free_slot = !(cur.lo >> 63) || (cur.lo & LV_TAG_MASK) != tag;
            if (!free_slot) break;
            const K128 mine{KW == 2 ? mine0 : (mine0 | KEY_LOCK), k[1]};
            const K128 old = cas128(slot, cur, mine);
            if (old.lo == cur.lo && old.hi == cur.hi) {
                if constexpr (KW > 2) {
#pragma unroll
                    for (int x = 2; x < KW; ++x) slot[x] = k[x];
                    __threadfence();
                    *(volatile uint64_t*)slot = mine0;   // unlock
                }
                *plen = i + 1;
                return 1;
            }
            cur = old;   // somebody else took the slot in this epoch: compare with what it wrote
        }
        if constexpr (KW == 2) {
            if ((cur.lo & ~LV_TAG_MASK) == k[0] && cur.hi == k[1]) { *plen = i + 1; return 0; }
        } else {
            if ((cur.lo & ~(LV_TAG_MASK | KEY_LOCK)) == k[0] && cur.hi == k[1]) {
                while (cur.lo & KEY_LOCK) cur.lo = ldcg64(slot);
                __threadfence();
                bool same = true;
#pragma unroll
                for (int x = 2; x < KW; ++x) same &= ldcg64(slot + x) == k[x];
                if (same) { *plen = i + 1; return 0; }
            }
        }
        idx = (idx + 1) & mask;
    }
    *plen = LV_MAX_PROBE;
    return -1;
}

__device__ __forceinline__ void lv_grid_barrier(unsigned long long* bar, unsigned long long& target, bool patient) {
    __syncthreads();
    if (threadIdx.x == 0) {
        target += gridDim.x;
        __threadfence();
        atomicAdd(bar, 1ull);
        unsigned ns = 20;
        while (ld_volatile(bar) < target) {
            if (patient) { __nanosleep(ns); if (ns < 400) ns += ns; }
        }
        __threadfence();
    }
    __syncthreads();
}

__host__ __device__ inline uint64_t lv_window(const LvParams& p, unsigned long long n_in, int boost) {
    uint64_t want = (uint64_t)n_in * p.slots_per_config;
    if (want < p.min_slots) want = p.min_slots;
    uint64_t s = 1;
    while (s < want) s <<= 1;
    for (int b = 0; b < boost && s < p.table_slots; ++b) s <<= 2;
    return s < p.table_slots ? s : p.table_slots;
}

// state transition after an attempt — evaluated identically by every thread that needs it
__device__ __forceinline__ void lv_advance(const LvParams& p, LvState& st, unsigned long long cnt, int stop, int cause, int retry) {
    st.attempt++;
    st.epoch = (st.epoch + 1) & 63;
    if (stop == 2) { st.stop = 2; st.cause = cause; return; }
    if (retry) {
        if (lv_window(p, st.n_in, st.boost) >= p.table_slots) { st.stop = 2; st.cause = JTB_CAUSE_TABLE_FULL; return; }
        st.boost++;
        return;   // same level, same input, larger window, new epoch
    }
    st.total += cnt;
    if (stop == 1) { st.stop = 1; return; }
    if (cnt == 0) { st.stop = 1; return; }                        // exhausted: the undecided shards are INVALID
    if (cnt > p.buf_cap) { st.stop = 2; st.cause = JTB_CAUSE_TABLE_FULL; return; }
    if (p.max_configs && st.total >= p.max_configs) { st.stop = 2; st.cause = JTB_CAUSE_BUDGET; return; }
    st.level++;
    st.n_in = cnt;
    st.in_idx ^= 1;
}

template <int MODEL, int KW, bool EAGER>
__global__ void __launch_bounds__(LV_THREADS, JTB_LV_CTAS) level_search_kernel(const LvParams p, const int neg_ok) {
    using L = EntryLayout<MODEL, KW>;
    constexpr int EW = L::EW;
    constexpr bool BAL = L::HAS_BAL;
    constexpr unsigned FULL = 0xffffffffu;
    using Scratch = LvScratch<KW, EW, BAL>;
    extern __shared__ __align__(16) unsigned char lv_smem[];
    __shared__ LvState s_state;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned lt_mask = (1u << lane) - 1;
    Scratch& S = reinterpret_cast<Scratch*>(lv_smem)[warp];
    LvCtrl* ctrl = p.ctrl;
    ExpandTables T;
    T.rows = p.rows; T.classes = p.classes; T.cls_inv_pos = p.cls_inv_pos; T.row_words = p.row_words; T.sum_off = p.sum_off;

    LvState st = p.init;
    unsigned long long bar_target = 0;
    unsigned long long my_probes = 0;
    int my_max_probe = 0;
    int wit_shard = -1, wit_rank = -1;   // lane-private filter for the witness atomicMax
    if (blockIdx.x == 0 && tid == 0) ctrl->t0 = globaltimer();

    // ---- one level attempt over the chunks [first, first + stride, ...) of the input array -------------------
    // Chunk = G configurations for one warp (G = 1, 2, .. 32: the smallest that gives every participating warp at most
    // one chunk, so a narrow level is spread over all warps and each has few children = few probe rounds).
    auto run_attempt = [&](const LvState& a, unsigned first_chunk, unsigned chunk_stride) {
        const uint64_t* in = p.buf[a.in_idx];
        uint64_t* out = p.buf[a.in_idx ^ 1];
        LvSlot* res = &ctrl->slot[a.attempt % 3];
        const uint64_t wmask = lv_window(p, a.n_in, a.boost) - 1;
        const uint64_t tag = (uint64_t)a.epoch << 56;
        unsigned G = 32;
        while (G > 1 && (a.n_in + (G >> 1) - 1) / (G >> 1) <= chunk_stride) G >>= 1;
        const unsigned n_chunks = (unsigned)((a.n_in + G - 1) / G);
        unsigned stg_head = 0, stg_tail = 0;   // warp-uniform
        auto flush = [&](unsigned n) {
            unsigned long long base = 0;
            if (lane == 0) base = atomicAdd(&res->cnt, (unsigned long long)n);
            base = __shfl_sync(FULL, base, 0);
            if (base + n <= p.buf_cap) {
                uint64_t* dst = out + base * EW;
                for (unsigned x = lane; x < n * EW; x += 32) {
                    const unsigned e = x / EW, k = x - e * EW;
                    dst[x] = S.stage[(stg_head + e) % LV_STAGE][k];
                }
            }   // else: cnt > buf_cap is seen by everyone after the barrier (TABLE_FULL)
            stg_head += n;
            __syncwarp();
        };
        for (unsigned chunk = first_chunk; chunk < n_chunks; chunk += chunk_stride) {
            // ---------------- phase 1: lane = configuration ------------------------------------------------
            const unsigned long long idx = (unsigned long long)chunk * G + lane;
            const bool have = (unsigned)lane < G && idx < a.n_in;
            uint64_t todo = 0, cls_todo = 0;
            {
                Expander<MODEL, KW, EAGER> X;
                X.todo = 0; X.rd_ok = 0; X.ncls = 0; X.cls_i = 0;
                X.fr_pos = 0; X.shard = 0; X.gj_end = 0; X.cls_base = 0; X.rslot = 0;
                if (have) {
                    const uint64_t* e = in + idx * EW;
#pragma unroll
                    for (int i = 0; i < KW; ++i) X.w[i] = ldcg64(e + i);
                    if constexpr (BAL) {
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const uint64_t v = ldcg64(e + KW + i);
                            X.bal[2 * i] = (int32_t)(uint32_t)v;
                            X.bal[2 * i + 1] = (int32_t)(uint32_t)(v >> 32);
                        }
                    }
                    const int shard = X.load_header(T);
                    const bool alive = !(p.n_shards > 1 && ld_volatile(&p.shard_found[shard]));
                    X.begin(T, alive);
                    todo = X.todo;
                    if (X.cls_i == 0) {   // not decided, not an exclusive eager read: crashed-op classes are candidates
                        for (int ci = 0; ci < X.ncls; ++ci) {
                            const int32_t* q = reinterpret_cast<const int32_t*>(T.classes + X.cls_base + ci);
                            const I4 b = ld_i4(q + 4);
                            const int shift = b.w & 0xff, width = b.w >> 8;
                            uint64_t field = 0;
#pragma unroll
                            for (int i = 1; i < KW; ++i) if (i == b.z) field = X.w[i];
                            const int count = (int)((field >> shift) & ((1ull << width) - 1));
                            if (count < b.y && ld_i32(T.cls_inv_pos + b.x + count) < X.fr_pos) cls_todo |= 1ull << ci;
                        }
                    }
#pragma unroll
                    for (int i = 0; i < KW; ++i) S.w[lane][i] = X.w[i];
                    if constexpr (BAL) {
#pragma unroll
                        for (int i = 0; i < 8; ++i) S.bal[lane][i] = X.bal[i];
                    }
                    S.rd_ok[lane] = X.rd_ok;
                    S.hdr[lane][0] = X.fr_pos; S.hdr[lane][1] = X.shard; S.hdr[lane][2] = X.gj_end;
                    S.hdr[lane][3] = X.cls_base; S.hdr[lane][4] = X.rslot; S.hdr[lane][5] = X.ncls;
                }
                S.todo[lane] = todo;
                S.cls_todo[lane] = cls_todo;
            }
            unsigned c = (unsigned)(__popcll(todo) + __popcll(cls_todo));
            unsigned incl = c;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned v = __shfl_up_sync(FULL, incl, o);
                if (lane >= o) incl += v;
            }
            S.start[lane] = incl - c;
            const unsigned total = __shfl_sync(FULL, incl, 31);
            if (lane == 0) S.start[32] = total;
            __syncwarp();
            // ---------------- phase 2: lane = child ------------------------------------------------------------
            for (unsigned g0 = 0; g0 < total; g0 += 32) {
                const unsigned g = g0 + lane;
                const bool act = g < total;
                Child<KW> ch;
                bool is_new = false;
                int owner = 0;
                if (act) {
                    int lo = 0, hi = 32;       // largest o with start[o] <= g
                    while (hi - lo > 1) {
                        const int mid = (lo + hi) >> 1;
                        if (S.start[mid] <= g) lo = mid; else hi = mid;
                    }
                    owner = lo;
                    const int k = (int)(g - S.start[owner]);
                    Expander<MODEL, KW, EAGER> Y;
#pragma unroll
                    for (int i = 0; i < KW; ++i) Y.w[i] = S.w[owner][i];
                    if constexpr (BAL) {
#pragma unroll
                        for (int i = 0; i < 8; ++i) Y.bal[i] = S.bal[owner][i];
                    }
                    Y.rd_ok = S.rd_ok[owner];
                    Y.gj = (int)((Y.w[0] >> 32) & XRANK_MASK);
                    Y.reg = (int32_t)(uint32_t)Y.w[0];
                    Y.row = T.rows + (size_t)Y.gj * T.row_words;
                    Y.fr_pos = S.hdr[owner][0]; Y.shard = S.hdr[owner][1]; Y.gj_end = S.hdr[owner][2];
                    Y.cls_base = S.hdr[owner][3]; Y.rslot = S.hdr[owner][4]; Y.ncls = S.hdr[owner][5];
                    const uint64_t otodo = S.todo[owner];
                    const int ns = __popcll(otodo);
                    bool ok;
                    int t_slot = 0;
                    if (k < ns) { t_slot = select64(otodo, k); ok = Y.child_slot(T, t_slot, neg_ok != 0, ch, true); }
                    else ok = Y.child_class(T, select64(S.cls_todo[owner], k - ns), neg_ok != 0, ch);
                    if (ok) {
                        if (ch.done) {
                            // every :ok op of the shard is linearized -> VALID
                            if (atomicExch(&p.shard_found[Y.shard], 1) == 0) {
                                if (atomicSub(&ctrl->n_undecided, 1) == 1) atomicCAS(&res->stop, 0, 1);
                            }
                        } else {
                            int plen;
                            const int r = lv_insert<KW>(p.table, wmask, ch.w, tag, &plen);
                            my_probes++;
                            my_max_probe = max(my_max_probe, plen);
                            if (r < 0) atomicExch(&res->retry, 1);
                            is_new = r == 1;
                            if constexpr (BAL) {
                                if (is_new && ch.d < 0) Y.load_transfer(t_slot, ch);   // only NEW children need the transfer
                            }
                            if (is_new && ch.cgj > Y.gj && (wit_shard != Y.shard || wit_rank < ch.cgj)) {
                                wit_shard = Y.shard; wit_rank = ch.cgj;
                                atomicMax(&p.shard_max_rank[Y.shard], ch.cgj);
                            }
                        }
                    }
                }
                // ---- stage the new children; append 32 at a time ----
                const unsigned newm = __ballot_sync(FULL, is_new);
                if (newm) {
                    if (is_new) {
                        uint64_t* e = S.stage[(stg_tail + __popc(newm & lt_mask)) % LV_STAGE];
#pragma unroll
                        for (int i = 0; i < KW; ++i) e[i] = ch.w[i];
                        if constexpr (BAL) {
                            int32_t b[8];
#pragma unroll
                            for (int i = 0; i < 8; ++i) b[i] = S.bal[owner][i];
                            if (ch.amt) {
#pragma unroll
                                for (int i = 0; i < 8; ++i) {
                                    if (i == ch.d) b[i] -= ch.amt;
                                    if (i == ch.c) b[i] += ch.amt;
                                }
                            }
#pragma unroll
                            for (int i = 0; i < 4; ++i) e[KW + i] = u64_of(b[2 * i], b[2 * i + 1]);
                        }
                    }
                    stg_tail += __popc(newm);
                    __syncwarp();
                    if (stg_tail - stg_head >= 32) flush(32);
                }
            }
            __syncwarp();   // the scratch of this chunk is dead
        }
        if (stg_tail != stg_head) flush(stg_tail - stg_head);
    };

    // ---- cooperative zero-fill of a grown window (the host cleared [0, zeroed_slots)) ---------------------------
    auto zero_fill = [&](uint64_t from, uint64_t to) {
        ulonglong2* t = reinterpret_cast<ulonglong2*>(p.table);
        const uint64_t n16 = (to - from) * KW / 2, off = from * KW / 2;
        for (uint64_t i = (uint64_t)blockIdx.x * LV_THREADS + tid; i < n16; i += (uint64_t)gridDim.x * LV_THREADS)
            t[off + i] = make_ulonglong2(0, 0);
    };

    for (;;) {
        if (st.stop) break;
        const bool narrow = st.n_in <= p.narrow_max;
        {   // a window larger than what is initialised: clear the new part first (wide attempts only; rare)
            const uint64_t win = lv_window(p, st.n_in, st.boost);
            if (win > st.zeroed) {
                zero_fill(st.zeroed, win);
                st.zeroed = win;
                lv_grid_barrier(&ctrl->bar, bar_target, false);
            }
        }
        if (!narrow) {
            if (blockIdx.x == 0 && tid == 0) {
                LvSlot* nx = &ctrl->slot[(st.attempt + 1) % 3];   // idle during this attempt: reset for the next one
                nx->cnt = 0; nx->stop = 0; nx->cause = 0; nx->retry = 0;
                if (p.time_budget_ns && globaltimer() - ctrl->t0 > p.time_budget_ns) {
                    LvSlot* res = &ctrl->slot[st.attempt % 3];
                    res->cause = JTB_CAUSE_BUDGET;
                    __threadfence();
                    atomicExch(&res->stop, 2);
                }
                if (st.n_in > ctrl->max_width) ctrl->max_width = st.n_in;
                const uint64_t win = lv_window(p, st.n_in, st.boost);
                if (win > ctrl->max_window) ctrl->max_window = win;
            }
            run_attempt(st, (unsigned)warp * gridDim.x + blockIdx.x, (unsigned)gridDim.x * LV_WARPS);
            lv_grid_barrier(&ctrl->bar, bar_target, false);
            const LvSlot* res = &ctrl->slot[st.attempt % 3];
            const unsigned long long cnt = ld_volatile(&res->cnt);
            const int stop = ld_volatile(&res->stop), cause = ld_volatile(&res->cause), retry = ld_volatile(&res->retry);
            lv_advance(p, st, cnt, stop, cause, retry);
            // wide -> narrow: CTA 0 is about to recycle the result slots on its own; everyone must have read this one
            if (!st.stop && st.n_in <= p.narrow_max) lv_grid_barrier(&ctrl->bar, bar_target, false);
        } else {
            if (blockIdx.x == 0) {
                // CTA 0 runs narrow levels on its own until the search widens, stops or ends
                for (;;) {
                    if (tid == 0) {
                        LvSlot* nx = &ctrl->slot[(st.attempt + 1) % 3];
                        nx->cnt = 0; nx->stop = 0; nx->cause = 0; nx->retry = 0;
                        if (p.time_budget_ns && (st.attempt & 63) == 0 && globaltimer() - ctrl->t0 > p.time_budget_ns) {
                            LvSlot* res = &ctrl->slot[st.attempt % 3];
                            res->cause = JTB_CAUSE_BUDGET;
                            res->stop = 2;
                        }
                        ctrl->narrow_levels++;
                    }
                    run_attempt(st, (unsigned)warp, (unsigned)LV_WARPS);
                    __syncthreads();
                    if (tid == 0) {
                        const LvSlot* res = &ctrl->slot[st.attempt % 3];
                        LvState nx = st;
                        lv_advance(p, nx, ld_volatile(&res->cnt), ld_volatile(&res->stop), ld_volatile(&res->cause),
                                   ld_volatile(&res->retry));
                        s_state = nx;
                    }
                    __syncthreads();
                    st = s_state;
                    if (st.stop || st.n_in > p.narrow_max || lv_window(p, st.n_in, st.boost) > st.zeroed) break;
                }
                if (tid == 0) { ctrl->pub = st; __threadfence(); }
            }
            lv_grid_barrier(&ctrl->bar, bar_target, blockIdx.x != 0);
            if (blockIdx.x != 0) {
                if (tid == 0) lv_load_state(s_state, &ctrl->pub);
                __syncthreads();
                st = s_state;
            }
        }
    }
    // ---- statistics ---------------------------------------------------------------------------------------
    for (int o = 16; o > 0; o >>= 1) {
        my_probes += __shfl_xor_sync(FULL, my_probes, o);
        my_max_probe = max(my_max_probe, __shfl_xor_sync(FULL, my_max_probe, o));
    }
    if (lane == 0) {
        if (my_probes) atomicAdd(&ctrl->probes, my_probes);
        atomicMax(&ctrl->max_probe_len, (unsigned long long)my_max_probe);
    }
    if (blockIdx.x == 0 && tid == 0) { ctrl->fin = st; ctrl->t1 = globaltimer(); }
}

}  // namespace jtb
