// jtb_wgl.cuh — Wing–Gong/Lowe linearizability search as a persistent sm_100a kernel.
//
// Replaces the hot loop of knossos.wgl/analysis behind jepsen.checker/linearizable (SURVEY.md §3.3,
// A.5): `cache.add((linearized BitSet, model))` + `model.step`.  B200-first formulation:
//
//  * A configuration is keyed EXACTLY by its window form
//        word0 = valid | global return rank of the first un-linearized :ok op | register state
//        word1 = bitmask over the (<=64) open-op slots at that return  + crashed-class counts
//        word2.. = more crashed-class counts (KW = 2, 4 or 8 words)
//    (every op that returned earlier is necessarily linearized; crashed ops of one (f, value) class
//    are consumed in invocation order so a count identifies the consumed set).
//  * One warp expands one configuration per step: lane t evaluates the op in open slot t
//    (model step inlined), the lane that linearizes the frontier op advances the frontier with
//    ballot/match arithmetic, every consistent child is probed/inserted in the global visited table
//    (16 B slots, ld.global.cg.v2.u64 probe + atom.cas.b128 insert), new children are pushed on the
//    CTA's shared-memory deque.  Depth-first order (LIFO) keeps the frontier deep.
//  * Work distribution: CTA-local deque in shared memory shared by its warps; oldest entries are
//    donated to a global LIFO pool when the deque is full or other CTAs are hungry; idle CTAs refill
//    from the pool.  Termination: no busy CTA and an empty pool (checked under the pool lock).
//  * Verdict: first config whose frontier passes the shard's last return => VALID (early exit);
//    exhaustion => INVALID with witness = furthest frontier rank reached (atomicMax), which is the
//    earliest :ok completion whose history prefix is not linearizable.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "jtb_prep.h"

namespace jtb {

struct __align__(16) K128 {
    uint64_t lo, hi;
};

struct Ctrl {
    // hot, read every step by one thread per CTA
    int stop;            // 0 run, 1 finished (all shards decided or exhausted), 2 aborted (cause)
    int n_hungry;        // CTAs with an empty deque
    int cause;           // JTB_CAUSE_* when stop == 2
    int n_busy;          // CTAs holding work
    int pool_lock;
    int n_undecided;     // shards not yet found VALID
    unsigned long long pool_top;   // entries in the global pool
    unsigned long long t0;         // %globaltimer at first CTA start
    // statistics (flushed at CTA exit)
    unsigned long long configs, probes, expansions, pool_pushes, pool_pops, idle_spins, max_probe_len;
};

struct WglParams {
    const int32_t* rows;
    const int4* ops;
    const int32_t* read_bal;
    const ClassRec* classes;
    const int32_t* cls_inv_pos;
    uint64_t* table;        // slots of KW 64-bit words
    uint64_t slot_mask;     // n_slots - 1
    uint64_t* pool;         // entries of EW words
    uint64_t pool_cap;      // entries
    Ctrl* ctrl;
    int* shard_found;       // [n_shards]
    int* shard_max_rank;    // [n_shards] furthest frontier reached (global rank)
    int row_words, S_pad, n_shards, max_nc;
    unsigned long long max_configs;   // stop (UNKNOWN) once this many configs were inserted
    int budget_cause;                 // JTB_CAUSE_BUDGET or JTB_CAUSE_TABLE_FULL (load guard)
    unsigned long long time_budget_ns;
    uint32_t deque_cap;     // power of two
};

constexpr uint64_t KEY_VALID = 1ull << 63;
constexpr uint64_t KEY_LOCK = 1ull << 62;   // only used by KW > 2 slots while their tail is written
constexpr int MAX_PROBE = 512;

// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ K128 ldcg128(const void* p) {
    K128 r;
    asm volatile("ld.global.cg.v2.u64 {%0, %1}, [%2];" : "=l"(r.lo), "=l"(r.hi) : "l"(p) : "memory");
    return r;
}
__device__ __forceinline__ uint64_t ldcg64(const void* p) {
    uint64_t r;
    asm volatile("ld.global.cg.u64 %0, [%1];" : "=l"(r) : "l"(p) : "memory");
    return r;
}
__device__ __forceinline__ K128 cas128(void* addr, K128 cmp, K128 val) {
    K128 old;
    asm volatile(
        "{\n\t"
        ".reg .b128 c, v, o;\n\t"
        "mov.b128 c, {%2, %3};\n\t"
        "mov.b128 v, {%4, %5};\n\t"
        "atom.relaxed.gpu.global.cas.b128 o, [%6], c, v;\n\t"
        "mov.b128 {%0, %1}, o;\n\t"
        "}\n"
        : "=l"(old.lo), "=l"(old.hi)
        : "l"(cmp.lo), "l"(cmp.hi), "l"(val.lo), "l"(val.hi), "l"(addr)
        : "memory");
    return old;
}
__device__ __forceinline__ int ld_volatile(const int* p) { return *(const volatile int*)p; }
__device__ __forceinline__ unsigned long long ld_volatile(const unsigned long long* p) {
    return *(const volatile unsigned long long*)p;
}
__device__ __forceinline__ unsigned long long globaltimer() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

__device__ __forceinline__ uint64_t mix64(uint64_t x) {
    x ^= x >> 33;
    x *= 0xff51afd7ed558ccdull;
    x ^= x >> 33;
    x *= 0xc4ceb9fe1a85ec53ull;
    x ^= x >> 33;
    return x;
}

template <int KW>
__device__ __forceinline__ uint64_t hash_key(const uint64_t (&k)[KW]) {
    uint64_t h = mix64(k[0] ^ 0x9E3779B97F4A7C15ull);
#pragma unroll
    for (int i = 1; i < KW; ++i) h = mix64(h ^ k[i] + 0x9E3779B97F4A7C15ull * (uint64_t)(i + 1));
    return h;
}

// Visited-table probe + insert.  Returns 1 = inserted (new config), 0 = already present,
// -1 = table exhausted.  *plen gets the number of slots inspected.
template <int KW>
__device__ __forceinline__ int table_insert(uint64_t* table, uint64_t slot_mask, const uint64_t (&k)[KW],
                                            int* plen) {
    uint64_t idx = hash_key<KW>(k) & slot_mask;
    for (int i = 0; i < MAX_PROBE; ++i) {
        uint64_t* slot = table + idx * KW;
        K128 cur = ldcg128(slot);
        if (cur.lo == 0 && cur.hi == 0) {
            K128 mine{KW == 2 ? k[0] : (k[0] | KEY_LOCK), k[1]};
            K128 old = cas128(slot, K128{0, 0}, mine);
            if (old.lo == 0 && old.hi == 0) {
                if constexpr (KW > 2) {
#pragma unroll
                    for (int w = 2; w < KW; ++w) slot[w] = k[w];
                    __threadfence();
                    *(volatile uint64_t*)slot = k[0];  // unlock
                }
                *plen = i + 1;
                return 1;
            }
            cur = old;
        }
        if constexpr (KW == 2) {
            if (cur.lo == k[0] && cur.hi == k[1]) { *plen = i + 1; return 0; }
        } else {
            if ((cur.lo & ~KEY_LOCK) == k[0] && cur.hi == k[1]) {
                while (cur.lo & KEY_LOCK) cur.lo = ldcg64(slot);
                __threadfence();
                bool same = true;
#pragma unroll
                for (int w = 2; w < KW; ++w) same &= ldcg64(slot + w) == k[w];
                if (same) { *plen = i + 1; return 0; }
            }
        }
        idx = (idx + 1) & slot_mask;
    }
    *plen = MAX_PROBE;
    return -1;
}

// ------------------------------------------------------------------------------------------------
// Model step, inlined per lane (knossos.model; bank: SURVEY §8(a) A7 from
// src/tigerbeetle/tests/ledger.clj:89-152).  `reg` is the register value (word0 low half),
// `bal` the 8 balances carried in the entry.
template <int MODEL>
__device__ __forceinline__ bool model_step(const int4 op, int32_t& reg, int32_t (&bal)[8],
                                           const int32_t* __restrict__ read_bal, bool neg_ok) {
    const int f = op.x & 0xff;
    if (op.x & OP_IMPOSSIBLE) return false;
    if constexpr (MODEL == JTB_MODEL_REGISTER || MODEL == JTB_MODEL_CAS_REGISTER) {
        if (f == JTB_F_READ) return op.y == JTB_NIL || op.y == reg;
        if (f == JTB_F_WRITE) { reg = op.y; return true; }
        if (reg != op.y) return false;  // cas
        reg = op.z;
        return true;
    } else if constexpr (MODEL == JTB_MODEL_BANK) {
        if (f == JTB_F_TRANSFER) {
            bool ok = true;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                if (i == op.z) { bal[i] -= op.y; ok &= neg_ok || bal[i] >= 0; }
                if (i == op.w) bal[i] += op.y;
            }
            return ok;
        }
        // read: every account present in the payload must match
        const int4* rb = reinterpret_cast<const int4*>(read_bal + (size_t)op.z * 8);
        const int4 lo = __ldg(rb), hi = __ldg(rb + 1);
        const int care = op.y;
        bool ok = true;
        ok &= !(care & 1) || bal[0] == lo.x;
        ok &= !(care & 2) || bal[1] == lo.y;
        ok &= !(care & 4) || bal[2] == lo.z;
        ok &= !(care & 8) || bal[3] == lo.w;
        ok &= !(care & 16) || bal[4] == hi.x;
        ok &= !(care & 32) || bal[5] == hi.y;
        ok &= !(care & 64) || bal[6] == hi.z;
        ok &= !(care & 128) || bal[7] == hi.w;
        return ok;
    } else {
        return false;
    }
}

// ------------------------------------------------------------------------------------------------
template <int MODEL, int KW>
struct EntryLayout {
    static constexpr bool HAS_BAL = MODEL == JTB_MODEL_BANK;
    static constexpr int EW = KW + (HAS_BAL ? 4 : 0);  // 64-bit words per deque/pool entry
};

constexpr int WGL_WARPS = 8;
constexpr int WGL_THREADS = WGL_WARPS * 32;

struct CtaShared {
    uint32_t top, bot;         // deque indices (monotonic, masked on use)
    uint32_t top_snap;
    int stop, hungry, idle;
    uint32_t xfer_n;
    unsigned long long xfer_off;
    unsigned long long wit_cache;  // (shard << 32 | furthest rank): filter for the witness atomicMax
    unsigned long long configs, probes, expansions, pool_pushes, pool_pops, idle_spins;
    int max_probe_len;
    int since_flush;
};

__device__ __forceinline__ void pool_lock(Ctrl* c) {
    while (atomicCAS(&c->pool_lock, 0, 1) != 0) __nanosleep(100);
    __threadfence();
}
__device__ __forceinline__ void pool_unlock(Ctrl* c) {
    __threadfence();
    atomicExch(&c->pool_lock, 0);
}

template <int MODEL, int KW>
__global__ void __launch_bounds__(WGL_THREADS, 4) wgl_search_kernel(const WglParams p, const int neg_ok,
                                                                    const int32_t init_reg) {
    using L = EntryLayout<MODEL, KW>;
    constexpr int EW = L::EW;
    extern __shared__ __align__(16) uint64_t s_deque[];  // deque_cap * EW words
    __shared__ CtaShared sh;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t cap_mask = p.deque_cap - 1;
    Ctrl* ctrl = p.ctrl;
    const int cand_rounds = p.S_pad / 32;
    const int cls_rounds = (p.max_nc + 31) / 32;
    const uint32_t worst_push = WGL_WARPS * 32 * (cand_rounds + cls_rounds);
    const uint32_t high = p.deque_cap - worst_push;

    if (tid == 0) {
        sh.top = sh.bot = 0;
        sh.stop = 0; sh.hungry = 0; sh.idle = 1;  // starts idle: not counted in n_busy
        sh.wit_cache = ~0ull;
        sh.configs = sh.probes = sh.expansions = sh.pool_pushes = sh.pool_pops = sh.idle_spins = 0;
        sh.max_probe_len = 0; sh.since_flush = 0;
        atomicAdd(&ctrl->n_hungry, 1);
        unsigned long long now = globaltimer();
        atomicCAS(&ctrl->t0, 0ull, now);
    }
    unsigned long long my_configs = 0, my_probes = 0;
    int my_steps = 0;
    int my_max_probe = 0;

    for (;;) {
        __syncthreads();  // (1) pushes of the previous step are complete
        if (tid == 0) {
            sh.stop = ld_volatile(&ctrl->stop);
            sh.hungry = ld_volatile(&ctrl->n_hungry);
            sh.top_snap = sh.top;
            if (++sh.since_flush >= 64) {  // budget checks, amortised
                sh.since_flush = 0;
                if (p.time_budget_ns && globaltimer() - ld_volatile(&ctrl->t0) > p.time_budget_ns) {
                    atomicCAS(&ctrl->cause, 0, JTB_CAUSE_BUDGET);
                    atomicExch(&ctrl->stop, 2);
                }
            }
        }
        __syncthreads();  // (2)
        if (sh.stop) break;
        uint32_t top = sh.top_snap;
        uint32_t size = top - sh.bot;

        if (size == 0) {
            // ---------------- acquire from the global pool (or detect termination) -------------
            if (tid == 0) {
                sh.xfer_n = 0;
                if (!sh.idle) {
                    sh.idle = 1;
                    atomicAdd(&ctrl->n_hungry, 1);
                    atomicSub(&ctrl->n_busy, 1);
                }
                unsigned long long pt = ld_volatile(&ctrl->pool_top);
                if (pt > 0) {
                    pool_lock(ctrl);
                    pt = ld_volatile(&ctrl->pool_top);
                    if (pt > 0) {
                        int hungry = max(1, ld_volatile(&ctrl->n_hungry));
                        unsigned long long take = pt / (unsigned)hungry;
                        take = take < 1 ? 1 : take > 4 * WGL_WARPS ? 4 * WGL_WARPS : take;
                        sh.xfer_n = (uint32_t)take;
                        sh.xfer_off = pt - take;
                        *(volatile unsigned long long*)&ctrl->pool_top = pt - take;
                        atomicAdd(&ctrl->n_busy, 1);
                        atomicSub(&ctrl->n_hungry, 1);
                        sh.idle = 0;
                        sh.pool_pops++;
                        // lock stays held until the entries are copied out
                    } else {
                        pool_unlock(ctrl);
                    }
                } else if (ld_volatile(&ctrl->n_busy) == 0) {
                    pool_lock(ctrl);
                    if (ld_volatile(&ctrl->pool_top) == 0 && ld_volatile(&ctrl->n_busy) == 0)
                        atomicCAS(&ctrl->stop, 0, 1);  // search space exhausted
                    pool_unlock(ctrl);
                } else {
                    sh.idle_spins++;
                    __nanosleep(500);
                }
            }
            __syncthreads();
            const uint32_t n = sh.xfer_n;
            if (n) {
                const uint64_t* src = p.pool + sh.xfer_off * EW;
                for (uint32_t i = tid; i < n * EW; i += WGL_THREADS) {
                    const uint32_t e = i / EW, w = i % EW;
                    s_deque[((top + e) & cap_mask) * EW + w] = ldcg64(src + i);
                }
                __syncthreads();
                if (tid == 0) {
                    pool_unlock(ctrl);
                    sh.top = top + n;
                }
            }
            continue;
        }

        if (size > high || (sh.hungry > 0 && size >= 2 * WGL_WARPS)) {
            // ---------------- donate the oldest entries to the global pool ----------------------
            const uint32_t n = size > high ? size - p.deque_cap / 4 : min(size / 2, 4u * WGL_WARPS * 4u);
            if (tid == 0) {
                pool_lock(ctrl);
                unsigned long long pt = ld_volatile(&ctrl->pool_top);
                sh.xfer_off = pt;
                sh.xfer_n = n;
                if (pt + n > p.pool_cap) {
                    sh.xfer_n = 0;
                    atomicCAS(&ctrl->cause, 0, JTB_CAUSE_BUDGET);
                    atomicExch(&ctrl->stop, 2);
                    pool_unlock(ctrl);
                }
            }
            __syncthreads();
            if (sh.xfer_n) {
                uint64_t* dst = p.pool + sh.xfer_off * EW;
                const uint32_t bot = sh.bot;
                for (uint32_t i = tid; i < n * EW; i += WGL_THREADS) {
                    const uint32_t e = i / EW, w = i % EW;
                    dst[i] = s_deque[((bot + e) & cap_mask) * EW + w];
                }
                __threadfence();
                __syncthreads();
                if (tid == 0) {
                    *(volatile unsigned long long*)&ctrl->pool_top = sh.xfer_off + n;
                    pool_unlock(ctrl);
                    sh.bot = bot + n;
                    sh.pool_pushes++;
                }
                size -= n;
            }
            __syncthreads();
            if (sh.xfer_n == 0) continue;  // aborted
        }

        // ---------------- pop: the top min(size, warps) entries, one per warp ----------------------
        const uint32_t take = min(size, (uint32_t)WGL_WARPS);
        uint64_t w[KW];
        int32_t pbal[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        const bool active = warp < (int)take;
        if (active) {
            const uint64_t* e = &s_deque[((top - 1 - warp) & cap_mask) * EW];
#pragma unroll
            for (int i = 0; i < KW; ++i) w[i] = e[i];
            if constexpr (L::HAS_BAL) {
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const uint64_t v = e[KW + i];
                    pbal[2 * i] = (int32_t)(uint32_t)v;
                    pbal[2 * i + 1] = (int32_t)(uint32_t)(v >> 32);
                }
            }
        }
        if (tid == 0) sh.top = top - take;
        __syncthreads();  // (3) popped entries are in registers; pushes may now reuse the space
        if (!active) continue;

        // ---------------- expand (warp-synchronous) -------------------------------------------------
        const int gj = (int)((w[0] >> 32) & 0x3fffffffu);
        const int32_t preg = (int32_t)(uint32_t)w[0];
        const int32_t* row = p.rows + (size_t)gj * p.row_words;
        const int32_t extra = __ldg(row + p.S_pad + (lane & 15));
        const int fr_pos = __shfl_sync(0xffffffffu, extra, 8);
        const int shard = __shfl_sync(0xffffffffu, extra, 9);
        const int gj_end = __shfl_sync(0xffffffffu, extra, 10);
        const int cls_base = __shfl_sync(0xffffffffu, extra, 11);
        const int ncls = __shfl_sync(0xffffffffu, extra, 12);
        const int rslot = __shfl_sync(0xffffffffu, extra, 13);
        if (p.n_shards > 1 && ld_volatile(&p.shard_found[shard])) continue;  // shard already VALID
        int n_new_total = 0, n_probe_total = 0;

        auto push_children = [&](bool is_new, const uint64_t (&cw)[KW], const int32_t (&cbal)[8]) {
            const unsigned newm = __ballot_sync(0xffffffffu, is_new);
            if (newm == 0) return;
            uint32_t base = 0;
            if (lane == 0) base = atomicAdd(&sh.top, (uint32_t)__popc(newm));
            base = __shfl_sync(0xffffffffu, base, 0);
            if (is_new) {
                uint64_t* e = &s_deque[((base + __popc(newm & ((1u << lane) - 1))) & cap_mask) * EW];
#pragma unroll
                for (int i = 0; i < KW; ++i) e[i] = cw[i];
                if constexpr (L::HAS_BAL) {
#pragma unroll
                    for (int i = 0; i < 4; ++i)
                        e[KW + i] = (uint64_t)(uint32_t)cbal[2 * i] | ((uint64_t)(uint32_t)cbal[2 * i + 1] << 32);
                }
            }
            n_new_total += __popc(newm);
        };

        // -- candidates: ops in the open slots
        for (int r = 0; r < cand_rounds; ++r) {
            const int t = r * 32 + lane;
            const int opid = __ldg(row + t);
            bool cand = opid >= 0 && !((w[1] >> t) & 1ull);
            int4 op = make_int4(OP_IMPOSSIBLE, 0, 0, 0);
            if (cand) op = __ldg(p.ops + opid);
            int32_t creg = preg;
            int32_t cbal[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) cbal[i] = pbal[i];
            bool ok = cand && model_step<MODEL>(op, creg, cbal, p.read_bal, neg_ok != 0);
            const bool is_front = t == rslot;
            uint64_t cw[KW];
#pragma unroll
            for (int i = 0; i < KW; ++i) cw[i] = w[i];
            int cgj = gj;
            // frontier advance, warp-cooperative, for the child that linearizes the frontier op
            const unsigned front_ok = __ballot_sync(0xffffffffu, ok && is_front);
            if (front_ok) {
                uint64_t m = w[1];
                int adv = 0;
                const int32_t* rw = row;
                int32_t ex = extra;
                for (;;) {
                    const int32_t word = __shfl_sync(0xffffffffu, ex, lane >> 2);
                    const int sl = (word >> (8 * (lane & 3))) & 0xff;
                    const bool setb = sl != 0xff && ((m >> sl) & 1ull);
                    const unsigned peers = __match_any_sync(0xffffffffu, sl);
                    const bool pass = setb && (peers & ((1u << lane) - 1)) == 0;
                    const unsigned pm = __ballot_sync(0xffffffffu, pass);
                    const int n = pm == 0xffffffffu ? 32 : __ffs(~pm) - 1;
                    const uint64_t clr = (lane < n) ? (1ull << sl) : 0ull;
                    const uint32_t clo = __reduce_or_sync(0xffffffffu, (uint32_t)clr);
                    const uint32_t chi = __reduce_or_sync(0xffffffffu, (uint32_t)(clr >> 32));
                    m &= ~((uint64_t)clo | ((uint64_t)chi << 32));
                    adv += n;
                    if (n < 32) break;
                    rw += (size_t)32 * p.row_words;
                    ex = __ldg(rw + p.S_pad + (lane & 15));
                }
                if (is_front) { cgj = gj + 1 + adv; cw[1] = m; }
            }
            if (ok && !is_front) cw[1] |= 1ull << t;
            cw[0] = KEY_VALID | ((uint64_t)(uint32_t)cgj << 32) |
                    ((MODEL == JTB_MODEL_BANK) ? 0ull : (uint64_t)(uint32_t)creg);
            int is_new = 0;
            if (ok) {
                if (cgj >= gj_end) {
                    // every :ok op of the shard is linearized -> VALID
                    if (atomicExch(&p.shard_found[shard], 1) == 0) {
                        if (atomicSub(&ctrl->n_undecided, 1) == 1) atomicCAS(&ctrl->stop, 0, 1);
                    }
                    ok = false;
                } else {
                    int plen;
                    const int res = table_insert<KW>(p.table, p.slot_mask, cw, &plen);
                    my_probes++;
                    my_max_probe = max(my_max_probe, plen);
                    if (res < 0) {
                        atomicCAS(&ctrl->cause, 0, JTB_CAUSE_TABLE_FULL);
                        atomicExch(&ctrl->stop, 2);
                    }
                    is_new = res > 0;
                    if (is_new && cgj > gj) {
                        // witness bookkeeping: furthest frontier reached in this shard
                        const unsigned long long wc = *(volatile unsigned long long*)&sh.wit_cache;
                        if ((int)(wc >> 32) != shard || (int)(uint32_t)wc < cgj) {
                            *(volatile unsigned long long*)&sh.wit_cache =
                                ((unsigned long long)(uint32_t)shard << 32) | (uint32_t)cgj;
                            atomicMax(&p.shard_max_rank[shard], cgj);
                        }
                    }
                }
            }
            n_probe_total += ok ? 1 : 0;
            push_children(is_new != 0, cw, cbal);
        }
        // -- candidates: next member of each crashed-op class
        for (int r = 0; r < cls_rounds; ++r) {
            const int c = r * 32 + lane;
            bool cand = c < ncls;
            struct { int first, n, word, shift_width; } cr = {0, 0, 1, 0};
            int4 cop = make_int4(OP_IMPOSSIBLE, 0, 0, 0);
            if (cand) {
                const int4* q = reinterpret_cast<const int4*>(p.classes + cls_base + c);
                const int4 b = __ldg(q + 1);
                cop = __ldg(q);
                cr.first = b.x; cr.n = b.y; cr.word = b.z; cr.shift_width = b.w;
            }
            const int shift = cr.shift_width & 0xff, width = cr.shift_width >> 8;
            uint64_t cw[KW];
            uint64_t field = 0;
#pragma unroll
            for (int i = 0; i < KW; ++i) { cw[i] = w[i]; if (i == cr.word) field = w[i]; }
            const int count = (int)((field >> shift) & ((1ull << width) - 1));
            cand = cand && count < cr.n;
            if (cand) cand = __ldg(p.cls_inv_pos + cr.first + count) < fr_pos;
            int32_t creg = preg;
            int32_t cbal[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) cbal[i] = pbal[i];
            const bool ok = cand && model_step<MODEL>(cop, creg, cbal, p.read_bal, neg_ok != 0);
#pragma unroll
            for (int i = 1; i < KW; ++i) if (i == cr.word) cw[i] += 1ull << shift;
            cw[0] = KEY_VALID | ((uint64_t)(uint32_t)gj << 32) |
                    ((MODEL == JTB_MODEL_BANK) ? 0ull : (uint64_t)(uint32_t)creg);
            int is_new = 0;
            if (ok) {
                int plen;
                const int res = table_insert<KW>(p.table, p.slot_mask, cw, &plen);
                my_probes++;
                my_max_probe = max(my_max_probe, plen);
                if (res < 0) {
                    atomicCAS(&ctrl->cause, 0, JTB_CAUSE_TABLE_FULL);
                    atomicExch(&ctrl->stop, 2);
                }
                is_new = res > 0;
            }
            push_children(is_new != 0, cw, cbal);
        }
        if (lane == 0) {
            my_configs += n_new_total;
            if (++my_steps >= 32 || my_configs >= 512) {
                // amortised global tally: budget (max_configs) and table-load guard
                my_steps = 0;
                const unsigned long long tot = atomicAdd(&ctrl->configs, my_configs) + my_configs;
                my_configs = 0;
                if (tot >= p.max_configs) {
                    atomicCAS(&ctrl->cause, 0, p.budget_cause);
                    atomicCAS(&ctrl->stop, 0, 2);
                }
            }
        }
        (void)init_reg;
    }
    // ---- flush statistics -------------------------------------------------------------------------
    for (int o = 16; o > 0; o >>= 1) {
        my_probes += __shfl_xor_sync(0xffffffffu, my_probes, o);
        my_max_probe = max(my_max_probe, __shfl_xor_sync(0xffffffffu, my_max_probe, o));
    }
    if (lane == 0) {
        atomicAdd(&ctrl->configs, my_configs);
        atomicAdd(&ctrl->probes, my_probes);
        atomicMax(&ctrl->max_probe_len, (unsigned long long)my_max_probe);
    }
    if (tid == 0) {
        atomicAdd(&ctrl->pool_pushes, sh.pool_pushes);
        atomicAdd(&ctrl->pool_pops, sh.pool_pops);
        atomicAdd(&ctrl->idle_spins, sh.idle_spins);
    }
}

}  // namespace jtb
