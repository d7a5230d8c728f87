// oracle_common.h — TEST INFRASTRUCTURE ONLY (CPU oracle).  Never linked into the product library.
//
// PARITY UNPINNED: the reference (nurturenature/jepsen-tigerbeetle) contains no golden vectors or
// known-answer tests for this path (test/tigerbeetle/core_test.clj:4-6 asserts nothing), and the
// algorithms live in un-vendored Maven deps (jepsen 0.2.8-SNAPSHOT, project.clj:8; knossos
// transitive, unpinned) that cannot run here (no JVM).  This oracle restates their PUBLISHED
// algorithms plus the in-tree Clojure (tests/ledger.clj:89-192,
// workloads/set_full.clj:51-75) and is pinned by (1) hand KATs from first principles
// and (2) four mutually independent implementations cross-checked under property tests.
//
// History preprocessing follows knossos.history/{complete,without-failures}:
//   * client ops only (process >= 0)                       tests/ledger.clj:94,204
//   * invoke paired with its completion by :process        knossos.history/pair-index
//   * :ok    -> op takes the completion's value (reads learn their value)
//   * :fail  -> op removed (did not happen)
//   * :info / never completed -> op stays open forever; crashed READS are dropped (no effect, no
//     constraint: a nil register read matches any state)
#pragma once
#include <algorithm>
#include <climits>
#include <cstdint>
#include <cstring>
#include <map>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <vector>

#include "../include/jtb_check.h"

namespace jtbo {

struct Op {
    int inv_pos = -1;          // position of the invoke among the shard's client events
    int ret_pos = INT_MAX;     // position of the :ok completion, INT_MAX if crashed
    int inv_index = -1;        // original :index of the invoke
    int ret_index = -1;        // original :index of the completion
    uint8_t f = 0;
    bool crashed = false;
    int32_t a = 0, b = 0, c = 0;
    const int32_t* pl = nullptr;  // payload (read values)
    int pl_len = -1;
    // derived
    std::vector<int> elems;    // set model: dense ids of the read's (deduplicated) payload
    bool impossible = false;   // op can never be linearized (e.g. set read of a never-added element)
    int cls = -1;              // crashed ops: equivalence class (same f,a,b,c)
    int rank_in_cls = 0;       // order of invocation inside the class
    int slot = -1;             // completed ops: interval-colouring slot (< 64)
};

struct State {
    int32_t reg = 0;
    int32_t bal[JTB_MAX_ACCOUNTS] = {0, 0, 0, 0, 0, 0, 0, 0};
    bool operator==(const State& o) const {
        return reg == o.reg && std::memcmp(bal, o.bal, sizeof bal) == 0;
    }
};

struct Shard {
    std::vector<Op> ops;             // in invocation order
    std::vector<int> rets;           // completed ops in return order
    int n_elems = 0;                 // set: number of distinct added elements
    int n_slots = 0;                 // max concurrently open completed ops
    int n_classes = 0;
    std::vector<std::vector<int>> cls_members;  // crashed-op classes, members in invocation order
    const jtb_model* model = nullptr;
    State init;
};

inline int acct_slot(const jtb_model* m, int32_t id) {
    for (int i = 0; i < m->n_accounts; ++i)
        if (m->account_ids[i] == id) return i;
    return -1;
}

inline Shard preprocess(const jtb_history* h, int s, const jtb_model* m) {
    Shard sh;
    sh.model = m;
    sh.init.reg = m->init_value;
    for (int i = 0; i < m->n_accounts && i < JTB_MAX_ACCOUNTS; ++i) sh.init.bal[i] = m->init_balance[i];
    std::unordered_map<int32_t, int> open;  // process -> op id
    std::vector<Op> all;
    std::vector<char> dropped;
    int pos = 0;
    for (int64_t e = h->shard_off[s]; e < h->shard_off[s + 1]; ++e) {
        if (h->process[e] < 0) continue;
        int32_t p = h->process[e];
        if (h->type[e] == JTB_T_INVOKE) {
            if (open.count(p)) throw std::runtime_error("process invoked twice without completing");
            Op o;
            o.inv_pos = pos++;
            o.inv_index = h->index[e];
            o.f = h->f[e];
            o.a = h->a[e]; o.b = h->b[e]; o.c = h->c[e];
            o.pl = h->payload + h->payload_off[e];
            o.pl_len = h->payload_len[e];
            open[p] = (int)all.size();
            all.push_back(o);
            dropped.push_back(0);
        } else {
            auto it = open.find(p);
            if (it == open.end()) throw std::runtime_error("completion without invocation");
            int id = it->second;
            open.erase(it);
            Op& o = all[id];
            ++pos;
            if (h->type[e] == JTB_T_OK) {
                o.ret_pos = pos - 1;
                o.ret_index = h->index[e];
                o.a = h->a[e]; o.b = h->b[e]; o.c = h->c[e];
                o.pl = h->payload + h->payload_off[e];
                o.pl_len = h->payload_len[e];
            } else if (h->type[e] == JTB_T_FAIL) {
                dropped[id] = 1;
            } else {  // info
                o.crashed = true;
                o.ret_index = h->index[e];
            }
        }
    }
    for (auto& kv : open) all[kv.second].crashed = true;
    for (size_t i = 0; i < all.size(); ++i) {
        if (dropped[i]) continue;
        if (all[i].crashed && all[i].f == JTB_F_READ) continue;  // crashed reads: drop
        sh.ops.push_back(all[i]);
    }
    // set model: dense element ids
    if (m->kind == JTB_MODEL_SET) {
        std::map<int32_t, int> dense;
        for (auto& o : sh.ops)
            if (o.f == JTB_F_ADD && !dense.count(o.a)) { int id = (int)dense.size(); dense[o.a] = id; }
        sh.n_elems = (int)dense.size();
        for (auto& o : sh.ops) {
            if (o.f == JTB_F_ADD) { o.a = dense[o.a]; continue; }
            if (o.pl_len < 0) { o.impossible = true; continue; }  // ok read of nil != any set
            std::vector<int32_t> v(o.pl, o.pl + o.pl_len);
            std::sort(v.begin(), v.end());
            v.erase(std::unique(v.begin(), v.end()), v.end());
            for (int32_t x : v) {
                auto it = dense.find(x);
                if (it == dense.end()) { o.impossible = true; break; }
                o.elems.push_back(it->second);
            }
        }
    }
    // return order
    for (int i = 0; i < (int)sh.ops.size(); ++i)
        if (!sh.ops[i].crashed) sh.rets.push_back(i);
    std::sort(sh.rets.begin(), sh.rets.end(),
              [&](int x, int y) { return sh.ops[x].ret_pos < sh.ops[y].ret_pos; });
    // interval colouring of completed ops (slot = smallest free)
    {
        std::vector<std::pair<int, int>> evs;  // (pos, op id) ; ret encoded as ~id
        for (int i = 0; i < (int)sh.ops.size(); ++i)
            if (!sh.ops[i].crashed) {
                evs.push_back({sh.ops[i].inv_pos, i});
                evs.push_back({sh.ops[i].ret_pos, ~i});
            }
        std::sort(evs.begin(), evs.end());
        std::vector<int> free_slots;
        int next = 0;
        for (auto& ev : evs) {
            if (ev.second >= 0) {
                int sl;
                if (free_slots.empty()) sl = next++;
                else {
                    auto it = std::min_element(free_slots.begin(), free_slots.end());
                    sl = *it; free_slots.erase(it);
                }
                sh.ops[ev.second].slot = sl;
            } else {
                free_slots.push_back(sh.ops[~ev.second].slot);
            }
        }
        sh.n_slots = next;
    }
    // crashed-op classes
    {
        std::map<std::tuple<int, int32_t, int32_t, int32_t>, int> cls;
        for (int i = 0; i < (int)sh.ops.size(); ++i) {
            Op& o = sh.ops[i];
            if (!o.crashed) continue;
            auto key = std::make_tuple((int)o.f, o.a, o.b, o.c);
            auto it = cls.find(key);
            if (it == cls.end()) {
                it = cls.emplace(key, (int)cls.size()).first;
                sh.cls_members.emplace_back();
            }
            o.cls = it->second;
            o.rank_in_cls = (int)sh.cls_members[o.cls].size();
            sh.cls_members[o.cls].push_back(i);
        }
        sh.n_classes = (int)cls.size();
    }
    return sh;
}

// Model step (knossos.model; bank: SURVEY §8(a) row A7 from tests/ledger.clj:89-152).
// Returns false when the op is inconsistent with the state.  For the set model `cnt` holds the
// multiplicity of each element among linearized adds and `distinct` the number of non-zero entries.
struct SetState {
    std::vector<int> cnt;
    int distinct = 0;
};

inline bool step(const Shard& sh, const Op& o, State& st, SetState* ss) {
    const jtb_model* m = sh.model;
    if (o.impossible) return false;
    switch (m->kind) {
    case JTB_MODEL_REGISTER:
    case JTB_MODEL_CAS_REGISTER:
        if (o.f == JTB_F_READ) return o.a == JTB_NIL || o.a == st.reg;
        if (o.f == JTB_F_WRITE) { st.reg = o.a; return true; }
        if (o.f == JTB_F_CAS && m->kind == JTB_MODEL_CAS_REGISTER) {
            if (st.reg != o.a) return false;
            st.reg = o.b;
            return true;
        }
        return false;
    case JTB_MODEL_SET:
        if (o.f == JTB_F_ADD) {
            if (ss->cnt[o.a]++ == 0) ss->distinct++;
            return true;
        }
        if (o.f == JTB_F_READ) {
            if ((int)o.elems.size() != ss->distinct) return false;
            for (int e : o.elems)
                if (ss->cnt[e] == 0) return false;
            return true;
        }
        return false;
    case JTB_MODEL_BANK:
        if (o.f == JTB_F_TRANSFER) {
            int d = acct_slot(m, o.b), c = acct_slot(m, o.c);
            if (d < 0 || c < 0) return false;
            st.bal[d] -= o.a;
            st.bal[c] += o.a;
            if (!m->negative_balances_ok && (st.bal[d] < 0 || st.bal[c] < 0)) return false;
            return true;
        }
        if (o.f == JTB_F_READ) {
            if (o.pl_len < 0) return false;
            for (int i = 0; i + 1 < o.pl_len; i += 2) {
                int sl = acct_slot(m, o.pl[i]);
                if (sl < 0) return false;                 // unknown key
                if (o.pl[i + 1] == JTB_NIL) return false;  // nil balance
                if (st.bal[sl] != o.pl[i + 1]) return false;
            }
            return true;
        }
        return false;
    }
    return false;
}

inline void unstep_set(const Op& o, SetState* ss) {
    if (o.f == JTB_F_ADD)
        if (--ss->cnt[o.a] == 0) ss->distinct--;
}

}  // namespace jtbo
