// hostwalk.cpp — TEST INFRASTRUCTURE (not part of libjtb_check.so, never loaded by the product).
//
// Drives the product's per-thread expansion core (jepsen_tigerbeetle_b200/csrc/jtb_expand.h — the code the search
// kernel inlines) and the product's host preparation (jtb_prep.cpp) on the CPU, with a std::unordered_set as the
// visited set, so that the candidate rules, the frontier advance, the eager-read rule and the crashed-class handling
// of the DEVICE code are compared with the oracle (verdict, witness, exhaustive configuration count) in the
// `-m "not gpu"` tier.  Depth-first, shard by shard; one configuration = one inserted key, exactly as on the device.
#include <cstring>
#include <string>
#include <unordered_set>
#include <vector>

#include "../../jepsen_tigerbeetle_b200/csrc/jtb_expand.h"

using namespace jtb;

namespace {

template <int KW>
struct KeyHash {
    size_t operator()(const std::string& s) const { return std::hash<std::string>()(s); }
};

template <int MODEL, int KW, bool EAGER>
void walk(const Prepared& P, const jtb_model* m, unsigned long long max_configs, int n_shards, int32_t* valid,
          int32_t* witness, int32_t* prev_ok, unsigned long long* configs_out) {
    const bool use_summary = false;   // the work-list kernels decide candidate reads cell by cell
    ExpandTables T{P.rows.data(), P.classes.data(), P.cls_inv_pos.data(), P.row_words, use_summary ? P.sum_off : 0};
    struct Entry { uint64_t w[KW]; int32_t bal[8]; };
    unsigned long long configs = 0;
    bool budget_hit = false;
    for (int s = 0; s < n_shards; ++s) {
        if (P.shard_cause[s]) { valid[s] = JTB_UNKNOWN; continue; }
        if (P.rank_base[s + 1] == P.rank_base[s]) { valid[s] = JTB_VALID; continue; }
        std::unordered_set<std::string> seen;
        std::vector<Entry> stack;
        Entry e0{};
        e0.w[0] = XKEY_VALID | ((uint64_t)(uint32_t)P.rank_base[s] << 32) |
                  ((MODEL == JTB_MODEL_BANK || MODEL == JTB_MODEL_SET) ? 0ull : (uint64_t)(uint32_t)m->init_value);
        for (int i = 0; i < 8; ++i) e0.bal[i] = m->init_balance[i];
        stack.push_back(e0);
        int max_rank = (int)P.rank_base[s];
        bool found = false;
        while (!stack.empty() && !found && !budget_hit) {
            Entry e = stack.back();
            stack.pop_back();
            Expander<MODEL, KW, EAGER> X;
            for (int i = 0; i < KW; ++i) X.w[i] = e.w[i];
            for (int i = 0; i < 8; ++i) X.bal[i] = e.bal[i];
            X.load_header(T);
            X.begin(T, true);
            Child<KW> ch;
            while (X.next(T, m->negative_balances_ok != 0, ch)) {
                if (ch.done) { found = true; break; }
                std::string key(reinterpret_cast<const char*>(ch.w), sizeof ch.w);
                if (!seen.insert(key).second) continue;
                ++configs;
                if (ch.cgj > max_rank) max_rank = ch.cgj;
                Entry c;
                for (int i = 0; i < KW; ++i) c.w[i] = ch.w[i];
                for (int i = 0; i < 8; ++i) c.bal[i] = e.bal[i];
                if (ch.amt) { c.bal[ch.d] -= ch.amt; c.bal[ch.c] += ch.amt; }
                stack.push_back(c);
                if (max_configs && configs >= max_configs) { budget_hit = true; break; }
            }
        }
        if (found) valid[s] = JTB_VALID;
        else if (budget_hit) valid[s] = JTB_UNKNOWN;
        else {
            valid[s] = JTB_INVALID;
            witness[s] = P.ret_index[max_rank];
            if (max_rank > P.rank_base[s]) prev_ok[s] = P.ret_index[max_rank - 1];
        }
    }
    *configs_out = configs;
}

// Level-synchronous twin of walk(): breadth-first by DEPTH (= number of linearized ops = rank + popcount(mask) + class
// counts; every move adds exactly one), with a visited set that only lives for ONE level — the device's level engine
// (csrc/jtb_bfs.cuh) relies on "two equal configurations always have equal depth", so a per-level set finds every
// duplicate.  All shards advance together, one op per level.  widths[l] (if given, cap entries) = configs at depth l+1.
template <int MODEL, int KW, bool EAGER>
void walk_bfs(const Prepared& P, const jtb_model* m, unsigned long long max_configs, int n_shards, int32_t* valid,
              int32_t* witness, int32_t* prev_ok, unsigned long long* configs_out, unsigned long long* widths, int cap,
              int* n_levels_out) {
    const bool use_summary = true;    // the level engine decides them from the rows' summary words
    ExpandTables T{P.rows.data(), P.classes.data(), P.cls_inv_pos.data(), P.row_words, use_summary ? P.sum_off : 0};
    struct Entry { uint64_t w[KW]; int32_t bal[8]; };
    std::vector<Entry> cur, nxt;
    std::vector<int> max_rank(n_shards);
    std::vector<char> found(n_shards, 0);
    for (int s = 0; s < n_shards; ++s) {
        max_rank[s] = (int)P.rank_base[s];
        if (P.shard_cause[s]) { valid[s] = JTB_UNKNOWN; found[s] = 2; continue; }
        if (P.rank_base[s + 1] == P.rank_base[s]) { valid[s] = JTB_VALID; found[s] = 2; continue; }
        Entry e0{};
        e0.w[0] = XKEY_VALID | ((uint64_t)(uint32_t)P.rank_base[s] << 32) |
                  ((MODEL == JTB_MODEL_BANK || MODEL == JTB_MODEL_SET) ? 0ull : (uint64_t)(uint32_t)m->init_value);
        for (int i = 0; i < 8; ++i) e0.bal[i] = m->init_balance[i];
        cur.push_back(e0);
    }
    unsigned long long configs = 0;
    bool budget_hit = false;
    int level = 0;
    while (!cur.empty() && !budget_hit) {
        std::unordered_set<std::string> seen;   // this level only
        nxt.clear();
        for (const Entry& e : cur) {
            Expander<MODEL, KW, EAGER> X;
            for (int i = 0; i < KW; ++i) X.w[i] = e.w[i];
            for (int i = 0; i < 8; ++i) X.bal[i] = e.bal[i];
            const int s = X.load_header(T);
            X.begin(T, !found[s]);
            Child<KW> ch;
            while (X.next(T, m->negative_balances_ok != 0, ch)) {
                if (ch.done) { found[s] = 1; break; }
                std::string key(reinterpret_cast<const char*>(ch.w), sizeof ch.w);
                if (!seen.insert(key).second) continue;
                ++configs;
                if (ch.cgj > max_rank[s]) max_rank[s] = ch.cgj;
                Entry c;
                for (int i = 0; i < KW; ++i) c.w[i] = ch.w[i];
                for (int i = 0; i < 8; ++i) c.bal[i] = e.bal[i];
                if (ch.amt) { c.bal[ch.d] -= ch.amt; c.bal[ch.c] += ch.amt; }
                nxt.push_back(c);
            }
            if (max_configs && configs >= max_configs) { budget_hit = true; break; }
        }
        if (widths && level < cap) widths[level] = nxt.size();
        ++level;
        cur.swap(nxt);
    }
    *n_levels_out = level;
    for (int s = 0; s < n_shards; ++s) {
        if (found[s] == 2) continue;
        if (found[s]) valid[s] = JTB_VALID;
        else if (budget_hit) valid[s] = JTB_UNKNOWN;
        else {
            valid[s] = JTB_INVALID;
            witness[s] = P.ret_index[max_rank[s]];
            if (max_rank[s] > P.rank_base[s]) prev_ok[s] = P.ret_index[max_rank[s] - 1];
        }
    }
    *configs_out = configs;
}

template <int MODEL, int KW>
void walk_e(bool eager, const Prepared& P, const jtb_model* m, unsigned long long mc, int ns, int32_t* v, int32_t* w,
            int32_t* pv, unsigned long long* c) {
    if (eager) walk<MODEL, KW, true>(P, m, mc, ns, v, w, pv, c);
    else walk<MODEL, KW, false>(P, m, mc, ns, v, w, pv, c);
}

template <int MODEL>
int walk_kw(int kw, bool eager, const Prepared& P, const jtb_model* m, unsigned long long mc, int ns, int32_t* v,
            int32_t* w, int32_t* pv, unsigned long long* c) {
    switch (kw) {
    case 2: walk_e<MODEL, 2>(eager, P, m, mc, ns, v, w, pv, c); return 0;
    case 4: walk_e<MODEL, 4>(eager, P, m, mc, ns, v, w, pv, c); return 0;
    case 8: walk_e<MODEL, 8>(eager, P, m, mc, ns, v, w, pv, c); return 0;
    }
    return -1;
}

template <int MODEL>
int bfs_kw(int kw, bool eager, const Prepared& P, const jtb_model* m, unsigned long long mc, int ns, int32_t* v,
           int32_t* w, int32_t* pv, unsigned long long* c, unsigned long long* widths, int cap, int* nl) {
    switch (kw) {
    case 2: if (eager) walk_bfs<MODEL, 2, true>(P, m, mc, ns, v, w, pv, c, widths, cap, nl);
            else walk_bfs<MODEL, 2, false>(P, m, mc, ns, v, w, pv, c, widths, cap, nl);
            return 0;
    case 4: if (eager) walk_bfs<MODEL, 4, true>(P, m, mc, ns, v, w, pv, c, widths, cap, nl);
            else walk_bfs<MODEL, 4, false>(P, m, mc, ns, v, w, pv, c, widths, cap, nl);
            return 0;
    case 8: if (eager) walk_bfs<MODEL, 8, true>(P, m, mc, ns, v, w, pv, c, widths, cap, nl);
            else walk_bfs<MODEL, 8, false>(P, m, mc, ns, v, w, pv, c, widths, cap, nl);
            return 0;
    }
    return -1;
}

}  // namespace

extern "C" int jtb_hostwalk_bfs(const jtb_history* h, const jtb_model* m, int eager, unsigned long long max_configs,
                                int32_t* valid, int32_t* witness, int32_t* prev_ok, unsigned long long* configs,
                                int32_t* key_words, unsigned long long* widths, int cap, int* n_levels) {
    Prepared P;
    if (!prepare(h, m, P)) return -3;
    for (int s = 0; s < h->n_shards; ++s) { valid[s] = JTB_UNKNOWN; witness[s] = prev_ok[s] = -1; }
    *key_words = P.key_words;
    const int ns = h->n_shards;
    switch (m->kind) {
    case JTB_MODEL_BANK: return bfs_kw<JTB_MODEL_BANK>(P.key_words, eager != 0, P, m, max_configs, ns, valid, witness, prev_ok, configs, widths, cap, n_levels);
    case JTB_MODEL_SET: return bfs_kw<JTB_MODEL_SET>(P.key_words, eager != 0, P, m, max_configs, ns, valid, witness, prev_ok, configs, widths, cap, n_levels);
    case JTB_MODEL_REGISTER:
    case JTB_MODEL_CAS_REGISTER:
        return bfs_kw<JTB_MODEL_CAS_REGISTER>(P.key_words, eager != 0, P, m, max_configs, ns, valid, witness, prev_ok, configs, widths, cap, n_levels);
    }
    return -2;
}

extern "C" int jtb_hostwalk(const jtb_history* h, const jtb_model* m, int eager, unsigned long long max_configs,
                            int32_t* valid, int32_t* witness, int32_t* prev_ok, unsigned long long* configs,
                            int32_t* key_words) {
    Prepared P;
    if (!prepare(h, m, P)) return -3;
    for (int s = 0; s < h->n_shards; ++s) { valid[s] = JTB_UNKNOWN; witness[s] = prev_ok[s] = -1; }
    *key_words = P.key_words;
    switch (m->kind) {
    case JTB_MODEL_BANK: return walk_kw<JTB_MODEL_BANK>(P.key_words, eager != 0, P, m, max_configs, h->n_shards, valid, witness, prev_ok, configs);
    case JTB_MODEL_SET: return walk_kw<JTB_MODEL_SET>(P.key_words, eager != 0, P, m, max_configs, h->n_shards, valid, witness, prev_ok, configs);
    case JTB_MODEL_REGISTER:
    case JTB_MODEL_CAS_REGISTER:
        return walk_kw<JTB_MODEL_CAS_REGISTER>(P.key_words, eager != 0, P, m, max_configs, h->n_shards, valid, witness, prev_ok, configs);
    }
    return -2;
}
