// scan_oracle.cpp — TEST INFRASTRUCTURE ONLY.  Sequential CPU restatements of the single-pass
// checkers on the hot path (see oracle_common.h for the parity statement):
//
//   jtbo_check_set_full     jepsen.checker/set-full, called at workloads/set_full.clj:157 with
//                           {:linearizable? true}; algorithm restated from SURVEY.md A.3
//                           (jepsen 0.2.x source, un-vendored — PARITY UNPINNED).
//   jtbo_check_bank_totals  the in-tree bank SI checker  tests/ledger.clj:116-192
//                           (check-op :127-152, err-badness :116-125, aggregation :154-192).
//
// Both are written as the literal event-by-event reduce the Clojure performs (maps keyed by element /
// process), deliberately NOT in the column-scan form the GPU kernels use.
#include <chrono>
#include <cmath>
#include <map>
#include <set>
#include <unordered_map>

#include "oracle_common.h"

namespace {

struct OpRef {
    int index = -1;
    int64_t time = 0;
    bool some = false;
};

struct Element {
    int32_t id;
    int order;  // creation order (for output)
    OpRef known, last_present, last_absent;
};

thread_local std::string g_err2;

}  // namespace

extern "C" {

const char* jtbo_scan_last_error(void) { return g_err2.c_str(); }

int jtbo_check_set_full(const jtb_history* h, int linearizable, jtb_setfull_out* out) {
    try {
        auto t0 = std::chrono::steady_clock::now();
        int64_t elem_cursor = 0;
        out->valid = JTB_VALID;
        out->n_failures = 0;
        if (out->elem_capacity > 0) out->elem_off[0] = 0;
        for (int s = 0; s < h->n_shards; ++s) {
            std::map<int32_t, Element> elements;       // keyed by element value
            std::vector<int32_t> creation;              // creation order of keys
            std::unordered_map<int32_t, OpRef> reads;   // process -> read invoke
            std::map<int32_t, int> dups;                // element -> max multiplicity
            int n_created = 0;
            for (int64_t e = h->shard_off[s]; e < h->shard_off[s + 1]; ++e) {
                if (h->process[e] < 0) continue;
                const int type = h->type[e], f = h->f[e];
                OpRef me{h->index[e], h->time_ns[e], true};
                if (f == JTB_F_ADD) {
                    const int32_t v = h->a[e];
                    if (type == JTB_T_INVOKE) {
                        Element el{v, n_created++, {}, {}, {}};
                        elements[v] = el;  // (assoc elements v (new element)) — re-add resets
                    } else if (type == JTB_T_OK) {
                        auto it = elements.find(v);
                        if (it != elements.end() && !it->second.known.some) it->second.known = me;
                    }
                } else if (f == JTB_F_READ) {
                    const int32_t p = h->process[e];
                    if (type == JTB_T_INVOKE) reads[p] = me;
                    else if (type == JTB_T_FAIL) reads.erase(p);
                    else if (type == JTB_T_OK) {
                        auto ri = reads.find(p);
                        if (ri == reads.end()) throw std::runtime_error("read :ok without invoke");
                        const OpRef inv = ri->second;
                        const int32_t* pl = h->payload + h->payload_off[e];
                        const int n = std::max(0, (int)h->payload_len[e]);
                        std::map<int32_t, int> mult;
                        for (int i = 0; i < n; ++i) mult[pl[i]]++;
                        for (auto& kv : mult)
                            if (kv.second > 1) dups[kv.first] = std::max(dups[kv.first], kv.second);
                        for (auto& kv : elements) {
                            Element& el = kv.second;
                            if (mult.count(kv.first)) {
                                if (!el.known.some) el.known = me;
                                if (!el.last_present.some || el.last_present.index < inv.index) el.last_present = inv;
                            } else {
                                if (!el.last_absent.some || el.last_absent.index < inv.index) el.last_absent = inv;
                            }
                        }
                    }
                }
            }
            // ---- results per element
            jtb_setfull_shard& r = out->shards[s];
            std::memset(&r, 0, sizeof r);
            std::vector<const Element*> ordered(elements.size());
            {
                std::vector<const Element*> tmp;
                for (auto& kv : elements) tmp.push_back(&kv.second);
                std::sort(tmp.begin(), tmp.end(), [](const Element* x, const Element* y) { return x->order < y->order; });
                ordered = tmp;
            }
            r.attempt_count = (int)elements.size();
            for (const Element* el : ordered) {
                const int lp = el->last_present.some ? el->last_present.index : -1;
                const int la = el->last_absent.some ? el->last_absent.index : -1;
                const bool stable = el->last_present.some && la < lp;
                const bool lost = el->known.some && el->last_absent.some && lp < la && el->known.index < la;
                int outcome = JTB_SF_NEVER_READ;
                int64_t lat_ms = 0;
                if (stable) {
                    const int64_t stable_time = el->last_absent.some ? el->last_absent.time + 1 : 0;
                    const int64_t d = std::max<int64_t>(0, stable_time - el->known.time);
                    lat_ms = (int64_t)((double)d / 1e6);  // (long (util/nanos->ms d))
                    outcome = JTB_SF_STABLE;
                    r.stable_count++;
                    if (lat_ms > 0) r.stale_count++;
                    r.stable_latency_max_ms = std::max(r.stable_latency_max_ms, lat_ms);
                } else if (lost) {
                    const int64_t lost_time = el->last_present.some ? el->last_present.time + 1 : 0;
                    const int64_t d = std::max<int64_t>(0, lost_time - el->known.time);
                    lat_ms = (int64_t)((double)d / 1e6);
                    outcome = JTB_SF_LOST;
                    r.lost_count++;
                    r.lost_latency_max_ms = std::max(r.lost_latency_max_ms, lat_ms);
                } else {
                    r.never_read_count++;
                }
                if (out->elem_capacity > 0) {
                    if (elem_cursor >= out->elem_capacity) throw std::runtime_error("elem_capacity too small");
                    out->elem_id[elem_cursor] = el->id;
                    out->elem_outcome[elem_cursor] = (uint8_t)outcome;
                    out->elem_latency_ms[elem_cursor] = lat_ms;
                    auto d = dups.find(el->id);
                    out->elem_dup_count[elem_cursor] = d == dups.end() ? 0 : d->second;
                    ++elem_cursor;
                }
            }
            r.duplicated_count = (int)dups.size();
            int valid;
            if (r.lost_count > 0) valid = JTB_INVALID;
            else if (r.stable_count == 0) valid = JTB_UNKNOWN;
            else if (linearizable && r.stale_count > 0) valid = JTB_INVALID;
            else valid = JTB_VALID;
            if (r.duplicated_count > 0) valid = JTB_INVALID;
            r.valid = valid;
            out->valid = std::max(out->valid, valid);
            out->n_failures += valid != JTB_VALID;
            if (out->elem_capacity > 0) out->elem_off[s + 1] = elem_cursor;
        }
        // ---- (read-all-invoked-adds), workloads/set_full.clj:51-75, literally: -------------------
        //   all-invoked-adds = set of :value over ops with :f :add and :type :invoke
        //   final-reads      = ops with :f :read, :type :ok, :final? true
        //   suspect          = final reads whose value does not contain every invoked add
        out->n_suspect = 0;
        out->raia_valid = JTB_VALID;
        int64_t missing_cursor = 0;
        if (out->suspect_capacity > 0) out->suspect_missing_off[0] = 0;
        for (int s = 0; s < h->n_shards; ++s) {
            std::set<int32_t> all_invoked_adds;
            for (int64_t e = h->shard_off[s]; e < h->shard_off[s + 1]; ++e)
                if (h->process[e] >= 0 && h->f[e] == JTB_F_ADD && h->type[e] == JTB_T_INVOKE) all_invoked_adds.insert(h->a[e]);
            int n_suspect_shard = 0;
            for (int64_t e = h->shard_off[s]; e < h->shard_off[s + 1]; ++e) {
                if (h->process[e] < 0 || h->f[e] != JTB_F_READ || h->type[e] != JTB_T_OK) continue;
                if (!h->flags || !(h->flags[e] & JTB_FLAG_FINAL)) continue;
                const int32_t* pl = h->payload + h->payload_off[e];
                std::set<int32_t> value(pl, pl + std::max(0, (int)h->payload_len[e]));
                std::vector<int32_t> missing;
                for (int32_t v : all_invoked_adds)
                    if (!value.count(v)) missing.push_back(v);
                if (missing.empty()) continue;
                ++n_suspect_shard;
                if (out->suspect_capacity > 0) {
                    if (out->n_suspect >= out->suspect_capacity || missing_cursor + (int64_t)missing.size() > out->missing_capacity)
                        throw std::runtime_error("suspect/missing capacity too small");
                    out->suspect_shard[out->n_suspect] = s;
                    out->suspect_index[out->n_suspect] = h->index[e];
                    for (int32_t v : missing) out->missing_ids[missing_cursor++] = v;
                    out->suspect_missing_off[out->n_suspect + 1] = missing_cursor;
                }
                out->n_suspect++;
            }
            out->shards[s].suspect_final_reads = n_suspect_shard;
            if (n_suspect_shard) out->raia_valid = JTB_INVALID;
        }
        out->seconds_total = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        out->seconds_kernel = out->seconds_total;
        return 0;
    } catch (const std::exception& e) {
        g_err2 = e.what();
        return -1;
    }
}

// tests/ledger.clj:116-125
static double err_badness(int type, int n_unexpected, int n_nil, int64_t total, int64_t total_amount,
                          int64_t neg_sum) {
    switch (type) {
    case JTB_BANK_UNEXPECTED_KEY: return n_unexpected;
    case JTB_BANK_NIL_BALANCE: return n_nil;
    case JTB_BANK_WRONG_TOTAL:
        // (Math/abs (float (/ (- total total-amount) total-amount))): division by a zero :total-amount throws in
        // Clojure (the caller flags reference_throws / :unknown); the :worst index is then ranked by |total - total-amount|.
        if (total_amount == 0) return std::fabs((double)(total - total_amount));
        return std::fabs((double)(float)((double)(total - total_amount) / (double)total_amount));
    case JTB_BANK_NEGATIVE_VALUE: return -(double)neg_sum;
    }
    return 0;
}

int jtbo_check_bank_totals(const jtb_history* h, const jtb_model* m, int64_t total_amount,
                           jtb_bank_result* out) {
    try {
        auto t0 = std::chrono::steady_clock::now();
        std::memset(out, 0, sizeof *out);
        out->first_error_index = -1;
        out->lowest_index = out->highest_index = -1;
        double worst_bad[5];
        for (int t = 0; t < 5; ++t) {
            out->first_index_by_type[t] = out->last_index_by_type[t] = out->worst_index_by_type[t] = -1;
            worst_bad[t] = -1e300;
        }
        bool have_total = false;
        for (int64_t e = 0; e < h->n_events; ++e) {
            if (h->process[e] < 0) continue;
            if (h->type[e] != JTB_T_OK || h->f[e] != JTB_F_READ) continue;
            out->read_count++;
            const int32_t* pl = h->payload + h->payload_off[e];
            const int n = std::max(0, (int)h->payload_len[e]);
            int n_unexpected = 0, n_nil = 0;
            int64_t total = 0, neg_sum = 0;
            bool any_neg = false;
            for (int i = 0; i + 1 < n; i += 2) {
                if (jtbo::acct_slot(m, pl[i]) < 0) n_unexpected++;
                if (pl[i + 1] == JTB_NIL) n_nil++;
                else {
                    total += pl[i + 1];
                    if (pl[i + 1] < 0) { any_neg = true; neg_sum += pl[i + 1]; }
                }
            }
            int type = JTB_BANK_OK;
            if (n_unexpected) type = JTB_BANK_UNEXPECTED_KEY;
            else if (n_nil) type = JTB_BANK_NIL_BALANCE;
            else if (total != total_amount) type = JTB_BANK_WRONG_TOTAL;
            else if (!m->negative_balances_ok && any_neg) type = JTB_BANK_NEGATIVE_VALUE;
            if (type == JTB_BANK_OK) continue;
            const int idx = h->index[e];
            out->error_count++;
            out->count_by_type[type]++;
            if (out->first_index_by_type[type] < 0) out->first_index_by_type[type] = idx;
            out->last_index_by_type[type] = idx;
            const double bad = err_badness(type, n_unexpected, n_nil, total, total_amount, neg_sum);
            if (bad > worst_bad[type]) { worst_bad[type] = bad; out->worst_index_by_type[type] = idx; }
            if (out->first_error_index < 0 || idx < out->first_error_index) {
                out->first_error_index = idx;
                out->first_error_type = type;
            }
            if (type == JTB_BANK_WRONG_TOTAL) {
                if (!have_total || total < out->lowest_total) { out->lowest_total = total; out->lowest_index = idx; }
                if (!have_total || total > out->highest_total) { out->highest_total = total; out->highest_index = idx; }
                have_total = true;
            }
        }
        out->valid = out->error_count ? JTB_INVALID : JTB_VALID;
        // tests/ledger.clj:122-123: (/ (- total total-amount) total-amount) with :total-amount 0 (the default, :356) is an
        // integer division by zero; util/max-by calls err-badness once an error type has >= 2 members, so the reference
        // checker THROWS there and jepsen's check-safe reports {:valid? :unknown}.
        if (total_amount == 0 && out->count_by_type[JTB_BANK_WRONG_TOTAL] >= 2) {
            out->reference_throws = 1;
            out->valid = JTB_UNKNOWN;
        }
        out->seconds_total = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        out->seconds_kernel = out->seconds_total;
        return 0;
    } catch (const std::exception& e) {
        g_err2 = e.what();
        return -1;
    }
}

}  // extern "C"
