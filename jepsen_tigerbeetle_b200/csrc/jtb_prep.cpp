// jtb_prep.cpp — see jtb_prep.h.  Host-side O(events) preparation: pairing, return ranks, open-op
// slots (interval colouring), crashed-op classes, frontier rows, per-model op resolution.
#include "jtb_prep.h"

#include <algorithm>
#include <climits>
#include <cstring>
#include <map>
#include <tuple>
#include <unordered_map>

namespace jtb {
namespace {

struct HOp {
    int inv_pos, ret_pos;      // ret_pos == INT_MAX for crashed ops
    int inv_ev, ret_ev;        // event numbers (absolute) of the invoke / completion, -1 if none
    bool crashed, dropped;
    int slot, rank;            // completed ops
    int cls;                   // crashed ops
};

struct ShardTmp {
    std::vector<HOp> ops;            // kept ops, invocation order
    std::vector<int> rets;           // completed op ids in return order
    int S = 0;
    std::vector<std::vector<int>> cls_members;
    std::vector<int> cls_rep;        // representative op of each class
    bool too_wide = false;
};

inline int acct_slot(const jtb_model* m, int32_t id) {
    for (int i = 0; i < m->n_accounts; ++i)
        if (m->account_ids[i] == id) return i;
    return -1;
}

// Resolve one op into its device record.  `ev` is the event whose value the op carries (the :ok
// completion for completed ops, the invoke for crashed ones).
OpRec resolve(const jtb_history* h, const jtb_model* m, int64_t ev, Prepared& out) {
    OpRec r{0, 0, 0, 0};
    const int f = h->f[ev];
    r.x = f;
    switch (m->kind) {
    case JTB_MODEL_REGISTER:
    case JTB_MODEL_CAS_REGISTER:
        r.y = h->a[ev];
        r.z = h->b[ev];
        if (f == JTB_F_CAS && m->kind != JTB_MODEL_CAS_REGISTER) r.x |= OP_IMPOSSIBLE;
        if (f != JTB_F_READ && f != JTB_F_WRITE && f != JTB_F_CAS) r.x |= OP_IMPOSSIBLE;
        break;
    case JTB_MODEL_BANK:
        if (f == JTB_F_TRANSFER) {
            const int d = acct_slot(m, h->b[ev]), c = acct_slot(m, h->c[ev]);
            r.y = h->a[ev];
            r.z = d;
            r.w = c;
            if (d < 0 || c < 0) r.x |= OP_IMPOSSIBLE;
        } else if (f == JTB_F_READ) {
            int32_t bal[JTB_MAX_ACCOUNTS] = {0};
            int care = 0;
            const int n = h->payload_len[ev];
            if (n < 0) r.x |= OP_IMPOSSIBLE;  // :ok read of nil
            const int32_t* pl = h->payload + h->payload_off[ev];
            for (int i = 0; i + 1 < n; i += 2) {
                const int sl = acct_slot(m, pl[i]);
                if (sl < 0 || pl[i + 1] == JTB_NIL) { r.x |= OP_IMPOSSIBLE; continue; }
                if ((care >> sl & 1) && bal[sl] != pl[i + 1]) r.x |= OP_IMPOSSIBLE;
                care |= 1 << sl;
                bal[sl] = pl[i + 1];
            }
            r.y = care;
            r.w = (int32_t)(out.read_bal.size() / JTB_MAX_ACCOUNTS);   // host-side index; replaced by the invocation position
            out.read_bal.insert(out.read_bal.end(), bal, bal + JTB_MAX_ACCOUNTS);
            if (care == (1 << m->n_accounts) - 1 && !(r.x & OP_IMPOSSIBLE)) {
                uint32_t hsh = 0;
                for (int i = 0; i < JTB_MAX_ACCOUNTS; ++i) hsh += (uint32_t)bal[i] * bank_hash_c(i);
                r.x |= OP_HASHED;
                r.z = (int32_t)hsh;
            }
        } else {
            r.x |= OP_IMPOSSIBLE;
        }
        break;
    case JTB_MODEL_SET:
        // adds: y = element value for now (dense id assigned by build_set_tables); reads are completed
        // by build_set_tables (per-(read, frontier) need/care masks)
        r.y = h->a[ev];
        if (f != JTB_F_ADD && f != JTB_F_READ) r.x |= OP_IMPOSSIBLE;
        break;
    default:
        r.x |= OP_IMPOSSIBLE;
    }
    return r;
}

// Grow-only set (knossos.model/set): a read of V is consistent iff the elements of the linearized adds
// are exactly V.  The linearized set is a function of the config, so the state needs no storage:
// for every read rho and every frontier rank g at which rho can be a candidate we precompute
//   need/care over key word 1  such that  rho is consistent  <=>  (word1 & care) == need.
// Adds that returned before rho was invoked are always linearized (must all be in V); adds invoked
// after rho returned never are; an overlapping add x must be linearized iff elem(x) in V:
//   rank(x) < g            -> linearized (fine if required, fatal if forbidden)
//   open at g (slot bit)   -> constrained bit
//   not yet invoked at g   -> not linearized (fatal if required)
// crashed adds contribute their 1-bit class count field the same way.
// Returns false when the shard needs something this encoding cannot express (an element of V whose only
// providers are two or more overlapping adds, or class fields outside word 1).
bool build_set_tables(const jtb_history* h, ShardTmp& t, const std::vector<int32_t>& gid, int64_t rank_base,
                      const std::vector<std::pair<int, int>>& field, const std::vector<int32_t>& fr_pos,
                      Prepared& out) {
    const int RW = out.row_words, SW = slot_words(JTB_MODEL_SET);
    // mark the inline record of op `gid_` impossible in every row it appears in: ranks [g0, j]
    auto mark_impossible = [&](int slot, int g0, int j) {
        for (int g = g0; g <= j; ++g) out.rows[(size_t)(rank_base + g) * RW + ROW_EXTRA + slot * SW] |= OP_IMPOSSIBLE;
    };
    struct Add { int32_t elem; bool crashed; int inv_pos, ret_pos, rank, slot, cls; };
    for (auto& f : field)
        if (f.first != 1) return false;
    for (auto& mem : t.cls_members)
        if (mem.size() > 1) return false;  // the same element crashed twice: multi-provider
    // adds per slot (completed, non-overlapping, in invocation order) and crashed adds in invocation order
    std::vector<std::vector<Add>> by_slot(64);
    std::vector<Add> crashed;
    std::vector<int> add_ret_sorted;       // ret_pos of completed adds, ascending
    std::vector<int32_t> elems;            // sorted unique element values of all adds
    for (int i = 0; i < (int)t.ops.size(); ++i) {
        const HOp& o = t.ops[i];
        const int64_t ev = o.crashed ? o.inv_ev : o.ret_ev;
        if (h->f[ev] != JTB_F_ADD) continue;
        Add a{h->a[ev], o.crashed, o.inv_pos, o.ret_pos, o.rank, o.slot, o.cls};
        elems.push_back(a.elem);
        if (o.crashed) crashed.push_back(a);
        else { by_slot[o.slot].push_back(a); add_ret_sorted.push_back(o.ret_pos); }
    }
    std::sort(add_ret_sorted.begin(), add_ret_sorted.end());
    std::sort(elems.begin(), elems.end());
    elems.erase(std::unique(elems.begin(), elems.end()), elems.end());
    // element value -> dense id: direct-address table when the value range is reasonably dense
    std::vector<int32_t> direct;
    const int64_t lo_e = elems.empty() ? 0 : elems.front(), hi_e = elems.empty() ? -1 : elems.back();
    if (!elems.empty() && hi_e - lo_e + 1 <= (int64_t)elems.size() * 256 + 4096) {
        direct.assign((size_t)(hi_e - lo_e + 1), -1);
        for (size_t k = 0; k < elems.size(); ++k) direct[(size_t)(elems[k] - lo_e)] = (int32_t)k;
    }
    auto elem_id = [&](int32_t e) -> int {
        if (!direct.empty()) return (e < lo_e || e > hi_e) ? -1 : direct[(size_t)(e - lo_e)];
        auto it = std::lower_bound(elems.begin(), elems.end(), e);
        return (it != elems.end() && *it == e) ? (int)(it - elems.begin()) : -1;
    };
    // per element: earliest return among its completed adds, and how many completed adds carry it
    std::vector<int> first_ret(elems.size(), INT_MAX), n_completed(elems.size(), 0);
    for (auto& v : by_slot)
        for (const Add& a : v) {
            const int id = elem_id(a.elem);
            first_ret[id] = std::min(first_ret[id], a.ret_pos);
            n_completed[id]++;
        }
    bool dup_completed = false;
    for (int c : n_completed) dup_completed |= c > 1;
    const int R = (int)t.rets.size();
    std::vector<int32_t> V;
    std::vector<const Add*> overlap;
    for (int j = 0; j < R; ++j) {
        const int i = t.rets[j];
        const HOp& ro = t.ops[i];
        if (h->f[ro.ret_ev] != JTB_F_READ) continue;
        OpRec& rec = out.ops[gid[i]];
        // first frontier rank at which rho is open: smallest g with fr_pos[g] > inv_pos(rho)
        const int g0 = (int)(std::upper_bound(fr_pos.begin(), fr_pos.end(), ro.inv_pos) - fr_pos.begin());
        const int n = h->payload_len[ro.ret_ev];
        if (n < 0) { rec.x |= OP_IMPOSSIBLE; mark_impossible(ro.slot, g0, j); continue; }
        const int32_t* pl = h->payload + h->payload_off[ro.ret_ev];
        V.assign(pl, pl + n);
        if (!std::is_sorted(V.begin(), V.end())) std::sort(V.begin(), V.end());
        V.erase(std::unique(V.begin(), V.end()), V.end());
        auto inV = [&](int32_t e) { return std::binary_search(V.begin(), V.end(), e); };
        // overlapping adds: per slot the run of ops with ret_pos > inv(rho) and inv_pos < ret(rho)
        overlap.clear();
        for (auto& v : by_slot) {
            auto it = std::partition_point(v.begin(), v.end(), [&](const Add& a) { return a.ret_pos < ro.inv_pos; });
            for (; it != v.end() && it->inv_pos < ro.ret_pos; ++it) overlap.push_back(&*it);
        }
        for (const Add& a : crashed) {
            if (a.inv_pos >= ro.ret_pos) break;
            overlap.push_back(&a);
        }
        // adds that returned before rho was invoked are always linearized: all of them must be in V
        const int n_before = (int)(std::lower_bound(add_ret_sorted.begin(), add_ret_sorted.end(), ro.inv_pos) -
                                   add_ret_sorted.begin());
        bool impossible = false;
        int before_in_V = 0;
        for (int32_t e : V) {
            const int id = elem_id(e);
            if (id < 0) { impossible = true; break; }        // nobody ever added it
            if (first_ret[id] < ro.inv_pos) {                 // covered by an add that already returned
                if (!dup_completed) before_in_V += 1;
                continue;
            }
            int providers = 0;
            for (const Add* x : overlap) providers += x->elem == e;
            if (providers == 0) { impossible = true; break; }
            if (providers > 1) return false;
        }
        if (!impossible) {
            if (dup_completed) {  // rare: count every before-add whose element is in V
                before_in_V = 0;
                for (auto& v : by_slot)
                    for (const Add& a : v)
                        if (a.ret_pos < ro.inv_pos && inV(a.elem)) ++before_in_V;
            }
            if (before_in_V != n_before) impossible = true;   // a returned add is missing from the read
        }
        if (impossible) { rec.x |= OP_IMPOSSIBLE; mark_impossible(ro.slot, g0, j); continue; }
        // classify the overlapping adds once
        struct Ov { const Add* x; bool want; };
        std::vector<Ov> ov;
        for (const Add* x : overlap) {
            const int id = elem_id(x->elem);
            const bool want = inV(x->elem);
            if (want && first_ret[id] < ro.inv_pos) continue;   // element already provided: unconstrained
            ov.push_back(Ov{x, want});
        }
        for (int g = g0; g <= j; ++g) {
            uint64_t need = 0, care = 0;
            bool feasible = true;
            for (const Ov& o2 : ov) {
                const Add* x = o2.x;
                if (x->crashed) {
                    const uint64_t bit = 1ull << (field[x->cls].second & 0xff);
                    if (x->inv_pos < fr_pos[g]) { care |= bit; if (o2.want) need |= bit; }
                    else if (o2.want) feasible = false;
                } else if (x->rank < g) {
                    if (!o2.want) feasible = false;                 // already linearized but not in V
                } else if (x->inv_pos < fr_pos[g]) {
                    const uint64_t bit = 1ull << x->slot;
                    care |= bit;
                    if (o2.want) need |= bit;
                } else if (o2.want) feasible = false;               // required but not invoked yet
            }
            if (!feasible) { need = ~0ull; care = 0; }              // (w & 0) == ~0 never holds
            int32_t* cell = &out.rows[(size_t)(rank_base + g) * RW + ROW_EXTRA + ro.slot * SW + 4];
            std::memcpy(cell, &need, 8);
            std::memcpy(cell + 2, &care, 8);
        }
    }
    return true;
}

}  // namespace

bool prepare(const jtb_history* h, const jtb_model* m, Prepared& out) {
    out = Prepared();
    out.model = m->kind;
    const int n_shards = h->n_shards;
    std::vector<ShardTmp> tmp(n_shards);
    out.shard_cause.assign(n_shards, JTB_CAUSE_NONE);
    out.max_classes.assign(n_shards, 0);
    if (m->kind == JTB_MODEL_BANK && (m->n_accounts < 1 || m->n_accounts > JTB_MAX_ACCOUNTS)) {
        out.error = "bank model needs 1..8 accounts";
        return false;
    }
    // ---- pass 1: pairing, ranks, slots, classes ---------------------------------------------
    int S_max = 1;
    for (int s = 0; s < n_shards; ++s) {
        ShardTmp& t = tmp[s];
        std::unordered_map<int32_t, int> open;
        std::vector<HOp> all;
        int pos = 0;
        for (int64_t e = h->shard_off[s]; e < h->shard_off[s + 1]; ++e) {
            const int32_t p = h->process[e];
            if (p < 0) continue;
            if (h->type[e] == JTB_T_INVOKE) {
                if (open.count(p)) { out.error = "process invoked twice without completing"; return false; }
                open[p] = (int)all.size();
                all.push_back(HOp{pos++, INT_MAX, (int)e, -1, false, false, -1, -1, -1});
            } else {
                auto it = open.find(p);
                if (it == open.end()) { out.error = "completion without invocation"; return false; }
                HOp& o = all[it->second];
                open.erase(it);
                o.ret_ev = (int)e;
                if (h->type[e] == JTB_T_OK) o.ret_pos = pos;
                else if (h->type[e] == JTB_T_FAIL) o.dropped = true;
                else o.crashed = true;
                ++pos;
            }
        }
        for (auto& kv : open) all[kv.second].crashed = true;
        for (auto& o : all) {
            if (o.dropped) continue;
            if (o.crashed && h->f[o.inv_ev] == JTB_F_READ) continue;
            t.ops.push_back(o);
        }
        // return ranks
        for (int i = 0; i < (int)t.ops.size(); ++i)
            if (!t.ops[i].crashed) t.rets.push_back(i);
        std::sort(t.rets.begin(), t.rets.end(),
                  [&](int x, int y) { return t.ops[x].ret_pos < t.ops[y].ret_pos; });
        for (int j = 0; j < (int)t.rets.size(); ++j) t.ops[t.rets[j]].rank = j;
        // slots: sweep in position order; ops are already sorted by inv_pos, returns by rank
        {
            uint64_t free_lo = ~0ull;  // bit set = slot free (slots 0..63)
            size_t ri = 0;
            int S = 0;
            for (int i = 0; i < (int)t.ops.size() && !t.too_wide; ++i) {
                HOp& o = t.ops[i];
                if (o.crashed) continue;
                while (ri < t.rets.size() && t.ops[t.rets[ri]].ret_pos < o.inv_pos) {
                    free_lo |= 1ull << t.ops[t.rets[ri]].slot;
                    ++ri;
                }
                if (!free_lo) { t.too_wide = true; break; }
                const int sl = __builtin_ctzll(free_lo);
                free_lo &= ~(1ull << sl);
                o.slot = sl;
                S = std::max(S, sl + 1);
            }
            t.S = S;
        }
        if (t.too_wide) {
            out.shard_cause[s] = JTB_CAUSE_TOO_WIDE;
            t.rets.clear();
            continue;
        }
        S_max = std::max(S_max, t.S);
        // crashed-op classes: same (f, a, b, c)
        std::map<std::tuple<int, int32_t, int32_t, int32_t>, int> cls;
        for (int i = 0; i < (int)t.ops.size(); ++i) {
            HOp& o = t.ops[i];
            if (!o.crashed) continue;
            const int64_t e = o.inv_ev;
            auto key = std::make_tuple((int)h->f[e], h->a[e], h->b[e], h->c[e]);
            auto it = cls.find(key);
            if (it == cls.end()) {
                it = cls.emplace(key, (int)cls.size()).first;
                t.cls_members.emplace_back();
                t.cls_rep.push_back(i);
            }
            o.cls = it->second;
            t.cls_members[o.cls].push_back(i);
        }
        out.max_classes[s] = (int)t.cls_members.size();
        out.max_nc = std::max(out.max_nc, out.max_classes[s]);
    }
    out.S_pad = S_max <= 32 ? 32 : 64;
    const int S_bits = S_max;
    // ---- pass 2: key layout (count fields of crashed-op classes) -----------------------------
    // word 0: valid | global rank | register state ; word 1: open-op mask in bits [0, S_bits),
    // class counts packed above it and into words 2.. (a field never straddles a word).
    int key_words = 2;
    std::vector<std::vector<std::pair<int, int>>> field(n_shards);  // (word, shift|width<<8)
    for (int s = 0; s < n_shards; ++s) {
        ShardTmp& t = tmp[s];
        if (out.shard_cause[s]) continue;
        int word = 1, bit = S_bits;
        for (auto& mem : t.cls_members) {
            int width = 1;
            while ((1 << width) <= (int)mem.size()) ++width;  // counts 0..n need ceil(log2(n+1)) bits
            if (bit + width > 64) { ++word; bit = 0; }
            field[s].push_back({word, bit | width << 8});
            bit += width;
        }
        int need = word + 1;
        int kw = need <= 2 ? 2 : need <= 4 ? 4 : need <= 8 ? 8 : 0;
        if (kw == 0) {
            out.shard_cause[s] = JTB_CAUSE_TOO_WIDE;
            t.rets.clear();
            field[s].clear();
            continue;
        }
        key_words = std::max(key_words, kw);
    }
    out.key_words = key_words;
    // ---- pass 3: tables -----------------------------------------------------------------------
    const int SW = slot_words(m->kind);
    out.sum_off = ROW_EXTRA + out.S_pad * SW;
    const int RW = out.sum_off + out.S_pad;
    out.row_words = RW;
    out.rank_base.assign(n_shards + 1, 0);
    for (int s = 0; s < n_shards; ++s) out.rank_base[s + 1] = out.rank_base[s] + (int64_t)tmp[s].rets.size();
    out.n_ranks = out.rank_base[n_shards];
    if (out.n_ranks >= (1ll << 29) - 64) { out.error = "history too large"; return false; }  // rank field: 29 bits
    out.rows.assign((size_t)out.n_ranks * RW, 0);
    out.ret_index.assign((size_t)out.n_ranks, -1);
    for (int s = 0; s < n_shards; ++s) {
        ShardTmp& t = tmp[s];
        const int R = (int)t.rets.size();
        if (R == 0) continue;
        const int64_t base = out.rank_base[s];
        const int32_t op_base = (int32_t)out.ops.size();
        // op records (completed ops only live in the table; crashed ones live in class records)
        std::vector<int32_t> gid(t.ops.size(), -1);
        std::vector<int32_t> op_inv_pos;   // by gid - op_base
        std::vector<int32_t> op_read_bal;  // by gid - op_base: bank reads' row in read_bal (resolve() left it in .w)
        for (int i = 0; i < (int)t.ops.size(); ++i) {
            if (t.ops[i].crashed) continue;
            gid[i] = (int32_t)out.ops.size();
            out.ops.push_back(resolve(h, m, t.ops[i].ret_ev, out));
            // ops carry their invocation position: the eager-read rule picks the EARLIEST-invoked consistent
            // read (a choice that does not depend on slot numbering; oracle: first in list order) and the
            // depth-first scouts order their candidates by it.  Bank transfers need .w for the credit slot:
            // theirs goes into the first (otherwise unused) payload word of the row cell.
            op_inv_pos.push_back(t.ops[i].inv_pos);
            op_read_bal.push_back(out.ops.back().w);
            const bool transfer = m->kind == JTB_MODEL_BANK && (out.ops.back().x & 0xff) == JTB_F_TRANSFER;
            if (!transfer) out.ops.back().w = t.ops[i].inv_pos;
        }
        // class records
        const int32_t cls_base = (int32_t)out.classes.size();
        for (size_t c = 0; c < t.cls_members.size(); ++c) {
            ClassRec cr;
            cr.op = resolve(h, m, t.ops[t.cls_rep[c]].inv_ev, out);
            cr.first = (int32_t)out.cls_inv_pos.size();
            cr.n = (int32_t)t.cls_members[c].size();
            cr.word = field[s][c].first;
            cr.shift_width = field[s][c].second;
            for (int i : t.cls_members[c]) out.cls_inv_pos.push_back(t.ops[i].inv_pos);
            out.classes.push_back(cr);
        }
        // rows: sweep invocations and returns in position order; slot contents are stored INLINE
        std::vector<int32_t> cur(out.S_pad, -1);
        size_t oi = 0;
        for (int j = 0; j < R; ++j) {
            const HOp& ro = t.ops[t.rets[j]];
            while (oi < t.ops.size() && t.ops[oi].inv_pos < ro.ret_pos) {
                if (!t.ops[oi].crashed) cur[t.ops[oi].slot] = gid[oi];
                ++oi;
            }
            int32_t* row = &out.rows[(size_t)(base + j) * RW];
            uint8_t* nxt = reinterpret_cast<uint8_t*>(row);
            for (int k = 0; k < 32; ++k)
                nxt[k] = (j + 1 + k < R) ? (uint8_t)t.ops[t.rets[j + 1 + k]].slot : (uint8_t)0xFF;
            row[8] = ro.ret_pos;
            row[9] = s;
            row[10] = (int32_t)(base + R);
            row[11] = cls_base;
            row[12] = (int32_t)t.cls_members.size();
            row[13] = ro.slot;
            for (int sl = 0; sl < out.S_pad; ++sl) {
                int32_t* cell = row + ROW_EXTRA + sl * SW;
                if (cur[sl] < 0) { cell[0] = OP_EMPTY; continue; }
                const OpRec& rec = out.ops[cur[sl]];
                cell[0] = rec.x; cell[1] = rec.y; cell[2] = rec.z; cell[3] = rec.w;
                if (m->kind == JTB_MODEL_BANK) {
                    if ((rec.x & 0xff) == JTB_F_READ)
                        std::memcpy(cell + 4, &out.read_bal[(size_t)op_read_bal[cur[sl] - op_base] * JTB_MAX_ACCOUNTS],
                                    8 * sizeof(int32_t));
                    else
                        cell[4] = op_inv_pos[cur[sl] - op_base];
                }
            }
            out.ret_index[base + j] = h->index[ro.ret_ev];
            cur[ro.slot] = -1;  // the op has returned
        }
        if (m->kind == JTB_MODEL_SET) {
            std::vector<int32_t> fr_pos(R);
            for (int j = 0; j < R; ++j) fr_pos[j] = t.ops[t.rets[j]].ret_pos;
            if (!build_set_tables(h, t, gid, base, field[s], fr_pos, out)) {
                // not expressible: the shard is reported UNKNOWN (too wide); neutralise its rows
                out.shard_cause[s] = JTB_CAUSE_TOO_WIDE;
            }
        }
        // slot masks of every row (after the set model has marked its impossible reads)
        for (int j = 0; j < R; ++j) {
            int32_t* row = &out.rows[(size_t)(base + j) * RW];
            uint64_t occ = 0, rd = 0, fast = 0;
            for (int sl = 0; sl < out.S_pad; ++sl) {
                const int32_t* cell = row + ROW_EXTRA + sl * SW;
                const int32_t x = cell[0];
                if (x < 0 || (x & OP_IMPOSSIBLE)) continue;
                occ |= 1ull << sl;
                if ((x & 0xff) != JTB_F_READ) continue;
                rd |= 1ull << sl;
                if (m->kind == JTB_MODEL_BANK) {
                    if (x & OP_HASHED) { fast |= 1ull << sl; row[out.sum_off + sl] = cell[2]; }
                } else if (m->kind != JTB_MODEL_SET) {
                    fast |= 1ull << sl;
                    row[out.sum_off + sl] = cell[1];
                }
            }
            std::memcpy(row + 14, &occ, 8);
            std::memcpy(row + 16, &rd, 8);
            std::memcpy(row + 18, &fast, 8);
        }
    }
    {
        unsigned long long open_sum = 0;
        for (int64_t g = 0; g < out.n_ranks; ++g) {
            uint64_t occ;
            std::memcpy(&occ, &out.rows[(size_t)g * RW + 14], 8);
            open_sum += (unsigned)__builtin_popcountll(occ);
        }
        out.mean_open = out.n_ranks ? (double)open_sum / (double)out.n_ranks : 0.0;
    }
    return true;
}

}  // namespace jtb
