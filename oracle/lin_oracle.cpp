// lin_oracle.cpp — TEST INFRASTRUCTURE ONLY.  CPU restatement of the linearizability analyses behind
// jepsen.checker/linearizable (see oracle_common.h for the parity statement).
//
// Four mutually independent deciders over the same preprocessed history:
//   ALGO_BRUTE   (0)  permutation enumeration, no memoisation, n <= ~9 ops          (ground truth)
//   ALGO_LINEAR  (1)  knossos.linear: event-ordered set of configs, just-in-time    
//   ALGO_WGL     (2)  knossos.wgl: Wing–Gong/Lowe DFS over a doubly linked entry list with
//                     lift/unlift and a cache of (linearized BitSet, model) — keyed literally on the
//                     full N-bit set, as Knossos does                                
//   ALGO_LEVEL   (4)  breadth-first by DEPTH (= number of linearized ops) over literal (linearized BitSet, model)
//                     configurations, with a visited set that is thrown away after every level: the rule the device's
//                     level engine relies on ("two equal configurations always have equal depth"), restated on the
//                     oracle's own data structures — independent of the product's keys, rows and expansion core
//   ALGO_LAZY_BANK (5) bank model only (negative balances allowed, reads of every account; crashed transfers enter only as
//                     part of a read's set): a REDUCED
//                     search — a transfer is linearized only when the frontier forces it or as part of the exact set
//                     that makes a read consistent — that visits 10^5 configurations where knossos.wgl visits 10^10.
//                     Not in Knossos; sound (argument in the struct's comment; verdict and witness equal to knossos.wgl on
//                     every random history tested).  Its purpose: an independent check of verdict and witness on the
//                     instances NO exhaustive CPU search can finish (bench headline: tau_think 0)
//   ALGO_WGL_COMPACT (3) the same DFS, cache keyed on the exact window form
//                     (first un-linearized return, mask of open ops, crashed bits, state).  This is
//                     the TIMED CPU BASELINE (bench.py cpu_baseline / --impl reference): Knossos'
//                     own cache would need N/8 bytes per config.
// Witness (":op") is defined canonically as the earliest :ok completion c such that the history
// prefix through c is not linearizable == the furthest return any branch was blocked at.
#include <chrono>
#include <functional>
#include <thread>
#include <atomic>
#include <array>
#include <set>

#include "oracle_common.h"

using namespace jtbo;

namespace {

enum { ALGO_BRUTE = 0, ALGO_LINEAR = 1, ALGO_WGL = 2, ALGO_WGL_COMPACT = 3, ALGO_LEVEL = 4, ALGO_LAZY_BANK = 5 };

struct Verdict {
    int valid = JTB_VALID;
    int witness_ret = -1;  // index into sh.rets
    int cause = JTB_CAUSE_NONE;
    uint64_t configs = 0, probes = 0;
};

// ------------------------------------------------------------------------------------------------
// open-addressing set of fixed-width keys (W 64-bit words); all-zero = empty (keys set bit 63 of w0)
class KeySet {
  public:
    explicit KeySet(int w) : W(w) { rehash(1 << 12); }
    bool insert(const uint64_t* k) {  // true if new
        if ((n + 1) * 10 > cap * 6) rehash(cap * 2);
        size_t i = hash(k) & (cap - 1);
        for (;;) {
            uint64_t* s = &slots[i * W];
            if (s[0] == 0) { std::memcpy(s, k, W * 8); ++n; return true; }
            if (std::memcmp(s, k, W * 8) == 0) return false;
            i = (i + 1) & (cap - 1);
        }
    }
    size_t size() const { return n; }
    template <class F>
    void for_each(F f) const {
        for (size_t i = 0; i < cap; ++i)
            if (slots[i * W]) f(&slots[i * W]);
    }

  private:
    int W;
    size_t cap = 0, n = 0;
    std::vector<uint64_t> slots;
    uint64_t hash(const uint64_t* k) const {
        uint64_t h = 0x9E3779B97F4A7C15ull;
        for (int i = 0; i < W; ++i) {
            h ^= k[i];
            h *= 0xFF51AFD7ED558CCDull;
            h ^= h >> 32;
        }
        return h;
    }
    void rehash(size_t ncap) {
        std::vector<uint64_t> old;
        old.swap(slots);
        size_t ocap = cap;
        cap = ncap;
        slots.assign(cap * W, 0);
        n = 0;
        for (size_t i = 0; i < ocap; ++i)
            if (old[i * W]) {
                size_t j = hash(&old[i * W]) & (cap - 1);
                while (slots[j * W]) j = (j + 1) & (cap - 1);
                std::memcpy(&slots[j * W], &old[i * W], W * 8);
                ++n;
            }
    }
};

// ------------------------------------------------------------------------------------------------
// ALGO_BRUTE
struct Brute {
    const Shard& sh;
    std::vector<int> ids;      // ops considered (invoked before the prefix end)
    std::vector<char> must;    // must be linearized (returned within the prefix)
    std::vector<int> perm;
    std::vector<char> used;
    bool found = false;
    explicit Brute(const Shard& s) : sh(s) {}

    bool order_ok() const {
        for (size_t i = 0; i < perm.size(); ++i)
            for (size_t j = i + 1; j < perm.size(); ++j) {
                // perm[i] linearized before perm[j]: illegal if perm[j] returned before perm[i] was invoked
                if (sh.ops[perm[j]].ret_pos < sh.ops[perm[i]].inv_pos) return false;
            }
        // every op left out must not be required; and a left-out op imposes nothing
        return true;
    }
    void rec(State st, SetState ss, int n_must_left, int prefix_end) {
        if (found) return;
        if (n_must_left == 0 && order_ok()) {
            // additionally: an included op must not be linearized after a *required* op that ... covered by order_ok
            found = true;
            return;
        }
        for (size_t k = 0; k < ids.size(); ++k) {
            if (used[k]) continue;
            const Op& o = sh.ops[ids[k]];
            bool legal = true;  // real-time order: o may not follow an op invoked after o returned
            for (int x : perm)
                if (o.ret_pos < sh.ops[x].inv_pos) { legal = false; break; }
            if (!legal) continue;
            State st2 = st;
            SetState ss2 = ss;
            if (!step(sh, o, st2, &ss2)) continue;
            used[k] = 1;
            perm.push_back(ids[k]);
            rec(st2, ss2, n_must_left - (must[k] ? 1 : 0), prefix_end);
            perm.pop_back();
            used[k] = 0;
            if (found) return;
        }
    }
    // is the prefix of the history up to and including position `end` linearizable?
    bool prefix_ok(int end) {
        ids.clear(); must.clear();
        int n_must = 0;
        for (int i = 0; i < (int)sh.ops.size(); ++i)
            if (sh.ops[i].inv_pos <= end) {
                ids.push_back(i);
                bool m = sh.ops[i].ret_pos <= end;
                must.push_back(m);
                n_must += m;
            }
        used.assign(ids.size(), 0);
        perm.clear();
        found = false;
        SetState ss;
        ss.cnt.assign(sh.n_elems, 0);
        rec(sh.init, ss, n_must, end);
        return found;
    }
    Verdict run() {
        Verdict v;
        if (sh.ops.size() > 12) throw std::runtime_error("brute force oracle limited to 12 ops");
        for (int j = 0; j < (int)sh.rets.size(); ++j)
            if (!prefix_ok(sh.ops[sh.rets[j]].ret_pos)) {
                v.valid = JTB_INVALID;
                v.witness_ret = j;
                return v;
            }
        return v;
    }
};

// ------------------------------------------------------------------------------------------------
// ALGO_LINEAR (knossos.linear, just-in-time)
struct Linear {
    const Shard& sh;
    uint64_t max_configs;
    explicit Linear(const Shard& s, uint64_t mc) : sh(s), max_configs(mc) {}
    struct Cfg {
        std::vector<int> ahead;  // sorted op ids linearized but not yet returned (crashed: forever)
        State st;
        bool operator<(const Cfg& o) const {
            if (ahead != o.ahead) return ahead < o.ahead;
            if (st.reg != o.st.reg) return st.reg < o.st.reg;
            return std::memcmp(st.bal, o.st.bal, sizeof st.bal) < 0;
        }
    };
    std::vector<int> returned_cnt;  // set model: element multiplicity among returned adds

    bool do_step(const Cfg& c, const Op& o, State& st2) const {
        st2 = c.st;
        if (sh.model->kind != JTB_MODEL_SET) return step(sh, o, st2, nullptr);
        if (o.f == JTB_F_ADD) return true;
        if (o.impossible) return false;
        // state = returned adds ∪ ahead adds
        std::set<int> extra;
        for (int x : c.ahead)
            if (sh.ops[x].f == JTB_F_ADD && returned_cnt[sh.ops[x].a] == 0) extra.insert(sh.ops[x].a);
        int distinct = 0;
        for (int e = 0; e < sh.n_elems; ++e) distinct += returned_cnt[e] > 0;
        distinct += (int)extra.size();
        if ((int)o.elems.size() != distinct) return false;
        for (int e : o.elems)
            if (returned_cnt[e] == 0 && !extra.count(e)) return false;
        return true;
    }

    Verdict run() {
        Verdict v;
        returned_cnt.assign(sh.n_elems, 0);
        // events
        std::vector<std::pair<int, int>> evs;  // (pos, op or ~op for return)
        for (int i = 0; i < (int)sh.ops.size(); ++i) {
            evs.push_back({sh.ops[i].inv_pos, i});
            if (!sh.ops[i].crashed) evs.push_back({sh.ops[i].ret_pos, ~i});
        }
        std::sort(evs.begin(), evs.end());
        std::set<Cfg> configs;
        configs.insert(Cfg{{}, sh.init});
        std::vector<int> pending;
        int ret_no = 0;
        for (auto& ev : evs) {
            if (ev.second >= 0) { pending.push_back(ev.second); continue; }
            int o = ~ev.second;
            std::set<Cfg> out, seen;
            for (const Cfg& c0 : configs) {
                std::vector<Cfg> stack{c0};
                while (!stack.empty()) {
                    Cfg c = stack.back();
                    stack.pop_back();
                    if (!seen.insert(c).second) continue;
                    v.configs++;
                    if (max_configs && v.configs > max_configs) {
                        v.valid = JTB_UNKNOWN; v.cause = JTB_CAUSE_BUDGET; return v;
                    }
                    auto it = std::lower_bound(c.ahead.begin(), c.ahead.end(), o);
                    if (it != c.ahead.end() && *it == o) {
                        Cfg r = c;
                        r.ahead.erase(r.ahead.begin() + (it - c.ahead.begin()));
                        out.insert(r);
                        continue;
                    }
                    for (int x : pending) {
                        if (std::binary_search(c.ahead.begin(), c.ahead.end(), x)) continue;
                        State st2;
                        if (!do_step(c, sh.ops[x], st2)) continue;
                        Cfg n;
                        n.st = st2;
                        if (x == o) {
                            n.ahead = c.ahead;
                            out.insert(n);
                        } else {
                            n.ahead = c.ahead;
                            n.ahead.insert(std::lower_bound(n.ahead.begin(), n.ahead.end(), x), x);
                            stack.push_back(n);
                        }
                    }
                }
            }
            if (out.empty()) {
                v.valid = JTB_INVALID;
                v.witness_ret = ret_no;
                return v;
            }
            // o has returned: it leaves `pending`; for the set model its element is now permanent
            pending.erase(std::find(pending.begin(), pending.end(), o));
            if (sh.model->kind == JTB_MODEL_SET && sh.ops[o].f == JTB_F_ADD) returned_cnt[sh.ops[o].a]++;
            configs.swap(out);
            ++ret_no;
        }
        return v;
    }
};

// ------------------------------------------------------------------------------------------------
// ALGO_WGL / ALGO_WGL_COMPACT  (knossos.wgl)
struct WGL {
    const Shard& sh;
    bool compact, canon, eager;
    uint64_t max_configs;
    // knossos :configs: when set (compact keys only) and the verdict is INVALID, receives every visited key whose
    // frontier is the witness (W words each), plus the initial configuration when the very first return is stuck
    std::vector<uint64_t>* final_keys = nullptr;
    int final_W = 0;
    // eager: "eager reads" reduction (NOT in Knossos; mirrors the device search so that exhaustive config
    // counts can be compared): a consistent read never changes the state, so when some candidate read is
    // consistent with the current config it is linearized immediately and EXCLUSIVELY (no sibling is
    // explored).  Sound for verdict and witness: any path from the config can be re-ordered to start with
    // that read (it stays enabled, states are unchanged, the frontier can only move forward).
    WGL(const Shard& s, bool compact_, bool canon_, uint64_t mc, bool eager_ = false)
        : sh(s), compact(compact_), canon(canon_), eager(eager_), max_configs(mc) {}

    Verdict run() {
        Verdict v;
        const int n = (int)sh.ops.size();
        const int n_ret = (int)sh.rets.size();
        if (n_ret == 0) return v;
        if (compact && sh.n_slots > 64) { v.valid = JTB_UNKNOWN; v.cause = JTB_CAUSE_TOO_WIDE; return v; }
        // entry list: node 2i = call of op i, 2i+1 = return of op i; node 2n = head sentinel
        const int HEAD = 2 * n;
        std::vector<int> nxt(2 * n + 1, -1), prv(2 * n + 1, -1);
        {
            std::vector<std::pair<int, int>> ent;
            for (int i = 0; i < n; ++i) {
                ent.push_back({sh.ops[i].inv_pos, 2 * i});
                if (!sh.ops[i].crashed) ent.push_back({sh.ops[i].ret_pos, 2 * i + 1});
            }
            std::sort(ent.begin(), ent.end());
            int last = HEAD;
            for (auto& e : ent) { nxt[last] = e.second; prv[e.second] = last; last = e.second; }
            nxt[last] = -1;
        }
        auto unlink = [&](int x) {
            nxt[prv[x]] = nxt[x];
            if (nxt[x] >= 0) prv[nxt[x]] = prv[x];
        };
        auto relink = [&](int x) {
            nxt[prv[x]] = x;
            if (nxt[x] >= 0) prv[nxt[x]] = x;
        };
        // crashed op numbering for the compact key
        std::vector<int> crash_no(n, -1);
        int n_crashed = 0;
        for (int i = 0; i < n; ++i)
            if (sh.ops[i].crashed) crash_no[i] = n_crashed++;
        const bool reg_model = sh.model->kind <= JTB_MODEL_CAS_REGISTER;
        const bool bank_model = sh.model->kind == JTB_MODEL_BANK;
        int W;
        if (compact) W = 2 + (n_crashed + 63) / 64;
        else W = 1 + (n + 63) / 64 + (bank_model ? 4 : 0);
        KeySet cache(W);
        std::vector<uint64_t> key(W);
        std::vector<uint64_t> lin((n + 63) / 64, 0), crashbits((n_crashed + 63) / 64 + 1, 0);
        std::vector<int> cls_count(sh.n_classes, 0);
        std::vector<int> ret_rank(n, -1);
        for (int j = 0; j < n_ret; ++j) ret_rank[sh.rets[j]] = j;

        struct Frame { int op; State st; int rj; uint64_t mask; bool exclusive; };
        std::vector<Frame> stack;
        State st = sh.init;
        SetState ss;
        ss.cnt.assign(sh.n_elems, 0);
        int rj = 0;          // first un-linearized return
        uint64_t mask = 0;   // slots of linearized completed ops that return after rets[rj]
        int max_rj = 0;
        int n_lin_completed = 0;

        auto make_key = [&](int rj2, uint64_t mask2, const State& st2) {
            if (compact) {
                key[0] = (1ull << 63) | ((uint64_t)(uint32_t)rj2 << 32) |
                         (reg_model ? (uint64_t)(uint32_t)st2.reg : 0ull);
                key[1] = mask2;
                for (int i = 0; i < W - 2; ++i) key[2 + i] = crashbits[i];
            } else {
                key[0] = (1ull << 63) | (reg_model ? (uint64_t)(uint32_t)st2.reg : 0ull);
                int nw = (n + 63) / 64;
                for (int i = 0; i < nw; ++i) key[1 + i] = lin[i];
                if (bank_model) std::memcpy(&key[1 + nw], st2.bal, 32);
            }
        };

        int entry = nxt[HEAD];
        bool fresh = true;        // just arrived at a config (eager-read scan pending)
        int only = -1;            // eager: the single call entry this config may linearize
        bool force_back = false;  // eager: the exclusive child is exhausted -> backtrack
        while (true) {
            if (n_lin_completed == n_ret) break;  // every :ok op linearized -> valid
            if (eager && fresh) {
                fresh = false;
                only = -1;
                for (int e = nxt[HEAD]; e >= 0 && !(e & 1); e = nxt[e]) {
                    const Op& o = sh.ops[e >> 1];
                    if (o.f != JTB_F_READ || o.crashed) continue;
                    State tmp = st;
                    if (step(sh, o, tmp, &ss)) { only = e; break; }
                }
                if (only >= 0) entry = only;
            }
            if (!force_back && entry >= 0 && !(entry & 1)) {
                // ---- call entry: try to linearize it
                const int i = entry >> 1;
                const Op& o = sh.ops[i];
                bool eligible = true;
                if (canon && o.crashed && cls_count[o.cls] != o.rank_in_cls) eligible = false;
                State st2 = st;
                bool ok = eligible && step(sh, o, st2, &ss);
                if (ok) {
                    // tentative new config
                    int rj2 = rj;
                    uint64_t mask2 = mask;
                    lin[i >> 6] |= 1ull << (i & 63);
                    if (o.crashed) crashbits[crash_no[i] >> 6] |= 1ull << (crash_no[i] & 63);
                    else if (sh.rets[rj] == i) {
                        ++rj2;
                        while (rj2 < n_ret && (lin[sh.rets[rj2] >> 6] >> (sh.rets[rj2] & 63) & 1)) {
                            mask2 &= ~(1ull << sh.ops[sh.rets[rj2]].slot);
                            ++rj2;
                        }
                    } else mask2 |= 1ull << o.slot;
                    make_key(rj2, mask2, st2);
                    v.probes++;
                    if (cache.insert(key.data())) {
                        v.configs++;
                        stack.push_back(Frame{i, st, rj, mask, only >= 0});
                        fresh = true;
                        st = st2; rj = rj2; mask = mask2;
                        if (rj > max_rj) max_rj = rj;
                        if (!o.crashed) { n_lin_completed++; unlink(2 * i + 1); }
                        else cls_count[o.cls]++;
                        unlink(2 * i);
                        entry = nxt[HEAD];
                        if (max_configs && v.configs >= max_configs && n_lin_completed != n_ret) {
                            v.valid = JTB_UNKNOWN; v.cause = JTB_CAUSE_BUDGET; return v;
                        }
                        continue;
                    }
                    // already seen: undo tentative marks
                    lin[i >> 6] &= ~(1ull << (i & 63));
                    if (o.crashed) crashbits[crash_no[i] >> 6] &= ~(1ull << (crash_no[i] & 63));
                    if (sh.model->kind == JTB_MODEL_SET) unstep_set(o, &ss);
                }
                if (only >= 0) force_back = true;  // the exclusive child was already visited
                else entry = nxt[entry];
            } else {
                // ---- return entry (or end of list): an un-linearized op has returned -> backtrack
                if (stack.empty()) {
                    v.valid = JTB_INVALID;
                    v.witness_ret = max_rj;
                    if (final_keys && compact) {
                        final_W = W;
                        if (max_rj == 0) {   // the initial configuration is never inserted
                            State s0 = sh.init;
                            std::fill(crashbits.begin(), crashbits.end(), 0);
                            make_key(0, 0, s0);
                            final_keys->insert(final_keys->end(), key.begin(), key.end());
                        }
                        cache.for_each([&](const uint64_t* k) {
                            if ((int)((k[0] >> 32) & 0x1fffffffu) == max_rj) final_keys->insert(final_keys->end(), k, k + W);
                        });
                    }
                    return v;
                }
                Frame fr = stack.back();
                stack.pop_back();
                const int i = fr.op;
                const Op& o = sh.ops[i];
                lin[i >> 6] &= ~(1ull << (i & 63));
                if (o.crashed) {
                    crashbits[crash_no[i] >> 6] &= ~(1ull << (crash_no[i] & 63));
                    cls_count[o.cls]--;
                } else {
                    n_lin_completed--;
                }
                if (sh.model->kind == JTB_MODEL_SET) unstep_set(o, &ss);
                st = fr.st; rj = fr.rj; mask = fr.mask;
                relink(2 * i);
                if (!o.crashed) relink(2 * i + 1);
                entry = nxt[2 * i];
                only = -1;
                force_back = fr.exclusive;  // an exclusive child has no siblings: keep backtracking
            }
        }
        return v;
    }
};

// ------------------------------------------------------------------------------------------------
// ALGO_LEVEL: level-synchronous search, visited set local to a level
struct LevelBFS {
    const Shard& sh;
    bool canon, eager;
    uint64_t max_configs;
    LevelBFS(const Shard& s, bool canon_, uint64_t mc, bool eager_) : sh(s), canon(canon_), eager(eager_), max_configs(mc) {}

    struct Cfg {
        std::vector<uint64_t> lin;
        State st;
        SetState ss;
        std::vector<int> cls_count;
        int rj = 0;   // first un-linearized return
    };

    Verdict run() {
        Verdict v;
        const int n = (int)sh.ops.size(), n_ret = (int)sh.rets.size();
        if (n_ret == 0) return v;
        const bool set_model = sh.model->kind == JTB_MODEL_SET;
        auto has = [](const Cfg& c, int i) { return (c.lin[i >> 6] >> (i & 63)) & 1ull; };
        Cfg c0;
        c0.lin.assign((n + 63) / 64, 0);
        c0.st = sh.init;
        c0.ss.cnt.assign(sh.n_elems, 0);
        c0.cls_count.assign(sh.n_classes, 0);
        std::vector<Cfg> cur{c0}, nxt;
        int max_rj = 0;
        while (!cur.empty()) {
            std::set<std::string> seen;   // this level only
            nxt.clear();
            for (const Cfg& c : cur) {
                const int frontier_pos = sh.ops[sh.rets[c.rj]].ret_pos;
                // candidates: un-linearized ops invoked before the first un-linearized return (ops are in invocation order)
                int only = -1;
                if (eager) {
                    for (int i = 0; i < n && sh.ops[i].inv_pos < frontier_pos; ++i) {
                        const Op& o = sh.ops[i];
                        if (has(c, i) || o.crashed || o.f != JTB_F_READ) continue;
                        State st2 = c.st;
                        SetState ss2 = c.ss;
                        if (step(sh, o, st2, &ss2)) { only = i; break; }
                    }
                }
                for (int i = 0; i < n && sh.ops[i].inv_pos < frontier_pos; ++i) {
                    if (only >= 0 && i != only) continue;
                    const Op& o = sh.ops[i];
                    if (has(c, i)) continue;
                    if (canon && o.crashed && c.cls_count[o.cls] != o.rank_in_cls) continue;
                    Cfg d;
                    d.st = c.st;
                    d.ss = c.ss;
                    if (!step(sh, o, d.st, &d.ss)) continue;
                    d.lin = c.lin;
                    d.lin[i >> 6] |= 1ull << (i & 63);
                    d.cls_count = c.cls_count;
                    if (o.crashed) d.cls_count[o.cls]++;
                    d.rj = c.rj;
                    while (d.rj < n_ret && has(d, sh.rets[d.rj])) ++d.rj;
                    std::string key(reinterpret_cast<const char*>(d.lin.data()), d.lin.size() * 8);
                    if (!set_model) key.append(reinterpret_cast<const char*>(&d.st), sizeof(State));
                    v.probes++;
                    if (!seen.insert(key).second) continue;
                    v.configs++;
                    if (d.rj > max_rj) max_rj = d.rj;
                    if (d.rj == n_ret) return v;   // every :ok op linearized -> valid
                    if (max_configs && v.configs >= max_configs) { v.valid = JTB_UNKNOWN; v.cause = JTB_CAUSE_BUDGET; return v; }
                    nxt.push_back(std::move(d));
                }
            }
            cur.swap(nxt);
        }
        v.valid = JTB_INVALID;
        v.witness_ret = max_rj;
        return v;
    }
};

// ------------------------------------------------------------------------------------------------
// ALGO_LAZY_BANK: "lazy transfers".  With negative balances allowed a transfer never fails and transfers commute, so the
// state after a set of transfers does not depend on their order, and a read (which reports every balance) pins the
// state.  Take any linearization and push every transfer as late as it can go: it stops either immediately before a read
// (together with the other transfers stuck there) or at its real-time deadline, i.e. when it is the frontier op.  So it is
// enough to search moves of two kinds:  (F) linearize the frontier op if it is a transfer;  (R) for a candidate read r,
// linearize a set D of pending transfers whose effects sum to exactly expected(r) - balances, then r (D = {}: the
// eager-read case, taken exclusively).  The furthest frontier reached (the witness) is preserved by the same pushing
// argument applied to the partial linearization that reaches it.  Configurations are keyed (frontier rank, open-slot mask).
struct LazyBank {
    const Shard& sh;
    uint64_t max_configs;
    LazyBank(const Shard& s, uint64_t mc) : sh(s), max_configs(mc) {}
    // crashed (:info) transfers are never forced; they only ever enter as part of a read's set D, members of one class
    // (same debit, credit, amount) in invocation order — so a configuration also records how many of each class it consumed
    struct Cfg { int rj; uint64_t mask; int32_t bal[JTB_MAX_ACCOUNTS]; std::vector<uint16_t> used; };

    Verdict run() {
        Verdict v;
        const jtb_model* m = sh.model;
        const int n = (int)sh.ops.size(), n_ret = (int)sh.rets.size();
        if (m->kind != JTB_MODEL_BANK || !m->negative_balances_ok) throw std::runtime_error("lazy-bank: bank model with negative balances allowed only");
        if (sh.n_slots > 64) { v.valid = JTB_UNKNOWN; v.cause = JTB_CAUSE_TOO_WIDE; return v; }
        for (const Op& o : sh.ops)
            if (!o.crashed && o.f == JTB_F_READ && !o.impossible && o.pl_len != 2 * m->n_accounts) throw std::runtime_error("lazy-bank: reads must cover every account");
        if (n_ret == 0) return v;
        std::vector<int> rank(n, -1);
        for (int j = 0; j < n_ret; ++j) rank[sh.rets[j]] = j;
        std::vector<std::vector<int>> open_at(n_ret);   // completed ops invoked before the j-th return, returning at or after it
        {
            std::vector<int> open;
            size_t oi = 0;
            for (int j = 0; j < n_ret; ++j) {
                const int fpos = sh.ops[sh.rets[j]].ret_pos;
                while (oi < sh.ops.size() && sh.ops[oi].inv_pos < fpos) { if (!sh.ops[oi].crashed) open.push_back((int)oi); ++oi; }
                open.erase(std::remove_if(open.begin(), open.end(), [&](int i) { return rank[i] < j; }), open.end());
                open_at[j] = open;
            }
        }
        auto usable = [&](const Op& o) { return o.f == JTB_F_TRANSFER && !o.impossible && acct_slot(m, o.b) >= 0 && acct_slot(m, o.c) >= 0; };
        auto eff = [&](const Op& o, int32_t* d) { d[acct_slot(m, o.b)] -= o.a; d[acct_slot(m, o.c)] += o.a; };
        const int ncls = sh.n_classes;
        std::set<std::string> seen;
        std::vector<Cfg> stack;
        Cfg c0{};
        for (int i = 0; i < JTB_MAX_ACCOUNTS; ++i) c0.bal[i] = sh.init.bal[i];
        c0.used.assign(ncls, 0);
        stack.push_back(c0);
        int max_rj = 0;
        bool found = false;
        auto push = [&](Cfg c) {
            while (c.rj < n_ret && ((c.mask >> sh.ops[sh.rets[c.rj]].slot) & 1ull)) { c.mask &= ~(1ull << sh.ops[sh.rets[c.rj]].slot); ++c.rj; }
            if (c.rj == n_ret) { found = true; return; }
            std::string key(reinterpret_cast<const char*>(&c.rj), 4);
            key.append(reinterpret_cast<const char*>(&c.mask), 8);
            key.append(reinterpret_cast<const char*>(c.used.data()), c.used.size() * 2);
            v.probes++;
            if (!seen.insert(key).second) return;
            v.configs++;
            max_rj = std::max(max_rj, c.rj);
            stack.push_back(std::move(c));
        };
        struct Item { int op; int cls; };   // cls >= 0: the next unused members of a crashed class, in invocation order
        while (!stack.empty() && !found) {
            if (max_configs && v.configs >= max_configs) { v.valid = JTB_UNKNOWN; v.cause = JTB_CAUSE_BUDGET; return v; }
            const Cfg c = stack.back();
            stack.pop_back();
            const int fpos = sh.ops[sh.rets[c.rj]].ret_pos;
            std::vector<Item> pend;
            std::vector<int> reads;
            for (int i : open_at[c.rj]) {
                if ((c.mask >> sh.ops[i].slot) & 1ull) continue;
                const Op& o = sh.ops[i];
                if (o.impossible) continue;
                if (o.f == JTB_F_TRANSFER) { if (usable(o)) pend.push_back(Item{i, -1}); }
                else if (o.f == JTB_F_READ) reads.push_back(i);
            }
            bool eager_done = false;   // a read that is consistent as things stand: the only child
            for (int r : reads) {
                State st;
                std::memcpy(st.bal, c.bal, sizeof st.bal);
                if (step(sh, sh.ops[r], st, nullptr)) { Cfg d = c; d.mask |= 1ull << sh.ops[r].slot; push(d); eager_done = true; break; }
            }
            if (eager_done) continue;
            const int f = sh.rets[c.rj];
            if (usable(sh.ops[f])) {   // (F)
                Cfg d = c;
                eff(sh.ops[f], d.bal);
                d.mask |= 1ull << sh.ops[f].slot;
                push(d);
            }
            for (int cl = 0; cl < ncls; ++cl) {   // crashed transfers invoked so far and not consumed yet, class by class
                const std::vector<int>& mem = sh.cls_members[cl];
                for (int k = c.used[cl]; k < (int)mem.size(); ++k) {
                    const Op& o = sh.ops[mem[k]];
                    if (o.inv_pos >= fpos || !usable(o)) break;
                    pend.push_back(Item{mem[k], cl});
                }
            }
            const int np = (int)pend.size();   // (R)
            std::vector<std::array<int32_t, JTB_MAX_ACCOUNTS>> pe(np), up(np + 1), dn(np + 1);
            for (int k = 0; k < np; ++k) { pe[k].fill(0); eff(sh.ops[pend[k].op], pe[k].data()); }
            up[np].fill(0); dn[np].fill(0);
            for (int k = np - 1; k >= 0; --k)
                for (int a = 0; a < JTB_MAX_ACCOUNTS; ++a) { up[k][a] = up[k + 1][a] + std::max(pe[k][a], 0); dn[k][a] = dn[k + 1][a] + std::min(pe[k][a], 0); }
            for (int r : reads) {
                const Op& ro = sh.ops[r];
                int32_t delta[JTB_MAX_ACCOUNTS] = {0}, want[JTB_MAX_ACCOUNTS];
                bool readable = true;
                std::memcpy(want, c.bal, sizeof want);
                for (int i = 0; i + 1 < ro.pl_len; i += 2) {
                    const int sl = acct_slot(m, ro.pl[i]);
                    if (sl < 0 || ro.pl[i + 1] == JTB_NIL) { readable = false; break; }
                    want[sl] = ro.pl[i + 1];
                }
                if (!readable) continue;
                for (int a = 0; a < JTB_MAX_ACCOUNTS; ++a) delta[a] = want[a] - c.bal[a];
                uint64_t dmask = 0;
                std::vector<uint16_t> used = c.used;
                std::function<void(int)> rec = [&](int k) {   // subsets of the pending transfers, bounded by what the rest can still move
                    bool zero = true;
                    for (int a = 0; a < JTB_MAX_ACCOUNTS; ++a) {
                        if (delta[a] > up[k][a] || delta[a] < dn[k][a]) return;
                        zero &= delta[a] == 0;
                    }
                    if (zero) {   // D found; a zero-sum superset is reachable from the child (its transfers are still pending there)
                        Cfg d = c;
                        d.mask |= dmask | (1ull << ro.slot);
                        d.used = used;
                        std::memcpy(d.bal, want, sizeof want);
                        push(d);
                        return;
                    }
                    if (k == np) return;
                    const Item it = pend[k];
                    // take it (a crashed member only when every earlier member of its class is taken: its turn)
                    if (it.cls < 0 || sh.ops[it.op].rank_in_cls == used[it.cls]) {
                        for (int a = 0; a < JTB_MAX_ACCOUNTS; ++a) delta[a] -= pe[k][a];
                        if (it.cls < 0) dmask |= 1ull << sh.ops[it.op].slot; else used[it.cls]++;
                        rec(k + 1);
                        if (it.cls < 0) dmask &= ~(1ull << sh.ops[it.op].slot); else used[it.cls]--;
                        for (int a = 0; a < JTB_MAX_ACCOUNTS; ++a) delta[a] += pe[k][a];
                    }
                    // skip it (and with it the later members of its class: they are interchangeable)
                    int k2 = k + 1;
                    if (it.cls >= 0) while (k2 < np && pend[k2].cls == it.cls) ++k2;
                    rec(k2);
                };
                rec(0);
            }
        }
        if (found) return v;
        v.valid = JTB_INVALID;
        v.witness_ret = max_rj;
        return v;
    }
};

void fill(const Shard& sh, const Verdict& v, jtb_lin_shard* out) {
    out->valid = v.valid;
    out->cause = v.cause;
    out->configs_explored = v.configs;
    out->probes = v.probes;
    out->witness_index = -1;
    out->previous_ok_index = -1;
    if (v.valid == JTB_INVALID && v.witness_ret >= 0) {
        out->witness_index = sh.ops[sh.rets[v.witness_ret]].ret_index;
        if (v.witness_ret > 0) out->previous_ok_index = sh.ops[sh.rets[v.witness_ret - 1]].ret_index;
    }
}

thread_local std::string g_err;

}  // namespace

extern "C" {

const char* jtbo_last_error(void) { return g_err.c_str(); }

// knossos :configs restated (the twin of jtb_final_configs, include/jtb_check.h): the configurations of an INVALID
// shard whose first un-linearized return is the witness, decoded and in the same canonical order.
// canon_info as in jtbo_check_linearizable (bit 0 is required: keys identify crashed ops by class counts there).
// Returns 0 and *n_total >= 0; *n_total = -1 when the shard is not INVALID.
int jtbo_final_configs(const jtb_history* h, const jtb_model* m, int canon_info, int32_t shard, jtb_final_config* out,
                       int32_t cap, int64_t* n_total) {
    try {
        Shard sh = preprocess(h, shard, m);
        WGL w(sh, true, true, 0, (canon_info & 2) != 0);
        std::vector<uint64_t> keys;
        w.final_keys = &keys;
        Verdict v = w.run();
        if (v.valid != JTB_INVALID) { *n_total = -1; return 0; }
        const int W = w.final_W;
        const int g = v.witness_ret;
        const Op& wit = sh.ops[sh.rets[g]];
        const bool bank = m->kind == JTB_MODEL_BANK;
        std::vector<int> rank(sh.ops.size(), -1), crashed_ops;
        for (int j = 0; j < (int)sh.rets.size(); ++j) rank[sh.rets[j]] = j;
        for (int i = 0; i < (int)sh.ops.size(); ++i)
            if (sh.ops[i].crashed) crashed_ops.push_back(i);
        auto apply = [&](int32_t* bal, const Op& o) {
            if (o.f != JTB_F_TRANSFER || o.impossible) return;
            const int d = acct_slot(m, o.b), c = acct_slot(m, o.c);
            if (d < 0 || c < 0) return;
            bal[d] -= o.a;
            bal[c] += o.a;
        };
        int32_t prefix[JTB_MAX_ACCOUNTS];
        for (int i = 0; i < JTB_MAX_ACCOUNTS; ++i) prefix[i] = sh.init.bal[i];
        if (bank)
            for (int j = 0; j < g; ++j) apply(prefix, sh.ops[sh.rets[j]]);
        std::vector<jtb_final_config> all(keys.size() / W);
        for (size_t k = 0; k < all.size(); ++k) {
            const uint64_t* key = &keys[k * W];
            jtb_final_config& c = all[k];
            std::memset(&c, 0, sizeof c);
            c.state = (bank || m->kind == JTB_MODEL_SET) ? 0 : (int32_t)(uint32_t)key[0];
            for (int i = 0; i < JTB_MAX_ACCOUNTS; ++i) c.balances[i] = bank ? prefix[i] : 0;
            for (int i = 0; i < (int)sh.ops.size(); ++i) {
                const Op& o = sh.ops[i];
                if (o.crashed || o.inv_pos > wit.ret_pos || rank[i] < g) continue;   // open at the witness' return
                if ((key[1] >> o.slot) & 1ull) {
                    c.linearized_open_index[c.n_linearized_open++] = o.inv_index;
                    if (bank) apply(c.balances, o);
                } else {
                    c.pending_index[c.n_pending++] = o.inv_index;
                }
            }
            std::sort(c.pending_index, c.pending_index + c.n_pending);
            std::sort(c.linearized_open_index, c.linearized_open_index + c.n_linearized_open);
            for (size_t b = 0; b < crashed_ops.size(); ++b)
                if ((key[2 + (b >> 6)] >> (b & 63)) & 1ull) {
                    c.n_crashed_linearized++;
                    if (bank) apply(c.balances, sh.ops[crashed_ops[b]]);
                }
        }
        std::sort(all.begin(), all.end(), [](const jtb_final_config& a, const jtb_final_config& b) {
            const int32_t* x = reinterpret_cast<const int32_t*>(&a);
            const int32_t* y = reinterpret_cast<const int32_t*>(&b);
            for (size_t i = 0; i < sizeof(jtb_final_config) / 4; ++i)
                if (x[i] != y[i]) return x[i] < y[i];
            return false;
        });
        *n_total = (int64_t)all.size();
        const size_t n_out = std::min<size_t>(all.size(), (size_t)std::max(cap, 0));
        if (n_out) std::memcpy(out, all.data(), n_out * sizeof(jtb_final_config));
        return 0;
    } catch (const std::exception& e) {
        g_err = e.what();
        return -1;
    }
}

// algo: 0 brute, 1 linear, 2 wgl (full-bitset cache), 3 wgl compact (timed baseline), 4 level (per-level visited set)
// canon_info: bit 0 = linearize crashed ops of one class (same f/value) in invocation order only;
//             bit 1 = eager reads (see WGL::eager; not part of Knossos)
// n_threads: shards are checked in parallel, one thread per shard at a time (independent/checker)
int jtbo_check_linearizable(const jtb_history* h, const jtb_model* m, int algo, uint64_t max_configs,
                            int canon_info, int n_threads, jtb_lin_shard* shards, jtb_lin_result* out) {
    try {
        auto t0 = std::chrono::steady_clock::now();
        std::atomic<int> next{0};
        std::atomic<bool> failed{false};
        std::string err;
        auto work = [&]() {
            for (;;) {
                int s = next.fetch_add(1);
                if (s >= h->n_shards) return;
                try {
                    Shard sh = preprocess(h, s, m);
                    Verdict v;
                    switch (algo) {
                    case ALGO_BRUTE: v = Brute(sh).run(); break;
                    case ALGO_LINEAR: v = Linear(sh, max_configs).run(); break;
                    case ALGO_WGL: v = WGL(sh, false, (canon_info & 1) != 0, max_configs, (canon_info & 2) != 0).run(); break;
                    case ALGO_WGL_COMPACT: v = WGL(sh, true, (canon_info & 1) != 0, max_configs, (canon_info & 2) != 0).run(); break;
                    case ALGO_LAZY_BANK: v = LazyBank(sh, max_configs).run(); break;
                    case ALGO_LEVEL: v = LevelBFS(sh, (canon_info & 1) != 0, max_configs, (canon_info & 2) != 0).run(); break;
                    default: throw std::runtime_error("unknown algo");
                    }
                    fill(sh, v, &shards[s]);
                } catch (const std::exception& e) {
                    if (!failed.exchange(true)) err = e.what();
                    return;
                }
            }
        };
        int nt = std::max(1, std::min(n_threads, (int)h->n_shards));
        if (nt == 1) work();
        else {
            std::vector<std::thread> th;
            for (int i = 0; i < nt; ++i) th.emplace_back(work);
            for (auto& t : th) t.join();
        }
        if (failed) { g_err = err; return -1; }
        std::memset(out, 0, sizeof *out);
        for (int s = 0; s < h->n_shards; ++s) {
            out->valid = std::max(out->valid, shards[s].valid);
            out->n_failures += shards[s].valid != JTB_VALID;
            out->configs_explored += shards[s].configs_explored;
            out->probes += shards[s].probes;
        }
        out->seconds_total = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        out->seconds_kernel = out->seconds_total;
        return 0;
    } catch (const std::exception& e) {
        g_err = e.what();
        return -1;
    }
}

}  // extern "C"
