// policy_sim.cpp — design experiment for the next round (NOT product code, not part of the library):
// how many configurations does a W-wide parallel search insert before it finds the linearization of a VALID
// history, as a function of the ORDER in which queued configurations are expanded?
//
// The shipped search kernel is depth-first inside a CTA and breadth-first across CTAs; on crash-heavy valid
// histories it drowns and the depth-first scouts rescue it slowly.  This simulates, on the CPU and
// with the oracle's preprocessing / model step, a synchronous machine that expands W configs per step chosen by a
// policy, with one global visited set — the shape of the GPU search with a different work queue:
//   fifo       breadth-first (the ticket ring alone)
//   lifo       global depth-first-ish stack (newest W)
//   rank       best-first by frontier rank (furthest first), newest first inside a rank
//   rank-crash best-first by rank, then FEWEST crashed ops consumed (dominating configs first)
//   bucket:S   what a GPU queue can afford: rank buckets of 2^S ranks, 16 sub-buckets by crashed ops consumed
//              relative to the bucket's running minimum (computed when the config is pushed), LIFO inside
// Register / cas-register only.  Build: g++ -O2 -std=c++17 -I../../include -I../../oracle policy_sim.cpp
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <queue>
#include <set>
#include <string>
#include <unordered_set>
#include <vector>

#include "oracle_common.h"

using namespace jtbo;

struct Cfg {
    int rj;
    uint64_t mask;
    int32_t reg;
    std::vector<uint8_t> cnt;   // crashed-class counts
    int crashed_used;
    uint64_t seq;
    int pb = 0, ps = 0;   // bucket / sub-bucket, fixed at push time (policy bucket:S)
    bool must_obs = false;   // JTB_SIM_LAZY=1: the last op was a crashed one, the next must observe the state
};
struct KeyHash {
    size_t operator()(const std::string& s) const { return std::hash<std::string>()(s); }
};
static std::string key_of(const Cfg& c) {
    std::string k(16 + c.cnt.size(), '\0');
    const int rj_flag = c.rj | (c.must_obs ? 1 << 30 : 0);
    std::memcpy(&k[0], &rj_flag, 4);
    std::memcpy(&k[4], &c.reg, 4);
    std::memcpy(&k[8], &c.mask, 8);
    if (!c.cnt.empty()) std::memcpy(&k[16], c.cnt.data(), c.cnt.size());
    return k;
}

struct Sim {
    Shard sh;
    std::vector<std::vector<int>> open_at;   // completed ops open at each return rank
    std::vector<int> rank;
    bool eager = true;
    bool lazy = std::getenv("JTB_SIM_LAZY") != nullptr;   // "lazy crashed ops" normal form

    explicit Sim(Shard s) : sh(std::move(s)) {
        const int R = (int)sh.rets.size();
        rank.assign(sh.ops.size(), -1);
        for (int j = 0; j < R; ++j) rank[sh.rets[j]] = j;
        open_at.resize(R);
        for (int j = 0; j < R; ++j) {
            const int rp = sh.ops[sh.rets[j]].ret_pos;
            for (int i = 0; i < (int)sh.ops.size(); ++i) {
                const Op& o = sh.ops[i];
                if (o.inv_pos > rp) break;   // ops are in invocation order
                if (!o.crashed && rank[i] >= j) open_at[j].push_back(i);
            }
        }
    }

    // children of c (eager-read rule as in the library); returns true when a child completes the history
    bool expand(const Cfg& c, std::vector<Cfg>& out) const {
        const int R = (int)sh.rets.size();
        const int rp = sh.ops[sh.rets[c.rj]].ret_pos;
        auto lin = [&](int i, Cfg n) {
            const Op& o = sh.ops[i];
            if (lazy && n.must_obs && o.f == JTB_F_WRITE) return false;
            State st; st.reg = n.reg;
            if (!step(sh, o, st, nullptr)) return false;
            if (lazy && o.crashed && st.reg == n.reg) return false;
            n.must_obs = lazy && o.crashed;
            n.reg = st.reg;
            if (o.crashed) { n.cnt[o.cls]++; n.crashed_used++; }
            else if (sh.rets[n.rj] == i) {
                ++n.rj;
                while (n.rj < R && ((n.mask >> sh.ops[sh.rets[n.rj]].slot) & 1ull)) {
                    n.mask &= ~(1ull << sh.ops[sh.rets[n.rj]].slot);
                    ++n.rj;
                }
            } else n.mask |= 1ull << o.slot;
            out.push_back(std::move(n));
            return true;
        };
        if (eager)
            for (int i : open_at[c.rj]) {
                const Op& o = sh.ops[i];
                if (o.f != JTB_F_READ || ((c.mask >> o.slot) & 1ull)) continue;
                State st; st.reg = c.reg;
                if (step(sh, o, st, nullptr)) { lin(i, c); return out.back().rj >= R; }
            }
        for (int i : open_at[c.rj])
            if (!((c.mask >> sh.ops[i].slot) & 1ull)) lin(i, c);
        for (int k = 0; k < sh.n_classes; ++k) {
            if (c.cnt[k] >= sh.cls_members[k].size()) continue;
            const int i = sh.cls_members[k][c.cnt[k]];
            if (sh.ops[i].inv_pos < rp) lin(i, c);
        }
        for (auto& n : out)
            if (n.rj >= R) return true;
        return false;
    }
};

int main(int argc, char** argv) {
    // input: a flattened single-shard register history dumped by policy_sim.py (binary, see there)
    if (argc < 5) { std::fprintf(stderr, "usage: policy_sim HISTORY.bin POLICY W MAX_CONFIGS\n"); return 2; }
    FILE* f = std::fopen(argv[1], "rb");
    if (!f) return 2;
    int64_t n; int32_t kind, init;
    if (std::fread(&n, 8, 1, f) != 1 || std::fread(&kind, 4, 1, f) != 1 || std::fread(&init, 4, 1, f) != 1) return 2;
    std::vector<uint8_t> type(n), ff(n), flags(n);
    std::vector<int32_t> process(n), index(n), a(n), b(n), c(n), plen(n, 0);
    std::vector<int64_t> time(n), poff(n, 0), shard_off{0, n}, key_ids{0};
    auto rd = [&](void* p, size_t sz) { return std::fread(p, sz, (size_t)n, f) == (size_t)n; };
    if (!rd(type.data(), 1) || !rd(ff.data(), 1) || !rd(process.data(), 4) || !rd(index.data(), 4) || !rd(a.data(), 4) ||
        !rd(b.data(), 4))
        return 2;
    std::fclose(f);
    jtb_history h{};
    h.n_events = n; h.type = type.data(); h.f = ff.data(); h.flags = flags.data(); h.process = process.data();
    h.index = index.data(); h.time_ns = time.data(); h.a = a.data(); h.b = b.data(); h.c = c.data();
    h.payload_off = poff.data(); h.payload_len = plen.data(); h.payload = nullptr; h.n_payload = 0;
    h.n_shards = 1; h.shard_off = shard_off.data(); h.key_ids = key_ids.data();
    jtb_model m{};
    m.kind = kind; m.init_value = init;
    Sim sim(preprocess(&h, 0, &m));
    const std::string policy = argv[2];
    const size_t W = (size_t)std::atoll(argv[3]);
    const uint64_t max_configs = (uint64_t)std::atoll(argv[4]);

    const bool bucketed = policy.rfind("bucket:", 0) == 0;
    const int shift = bucketed ? std::atoi(policy.c_str() + 7) : 0;
    std::vector<int> min_crashed((sim.sh.rets.size() >> shift) + 2, 1 << 30);
    auto prio = [&](const Cfg& x) -> std::tuple<long long, long long, long long> {   // larger = expanded earlier
        if (bucketed) return {x.pb, -x.ps, (long long)x.seq};
        if (policy == "fifo") return {0, 0, -(long long)x.seq};
        if (policy == "lifo") return {0, 0, (long long)x.seq};
        if (policy == "rank") return {x.rj, 0, (long long)x.seq};
        return {x.rj, -x.crashed_used, (long long)x.seq};                        // rank-crash
    };
    auto cmp = [&](const Cfg& x, const Cfg& y) { return prio(x) < prio(y); };
    std::priority_queue<Cfg, std::vector<Cfg>, decltype(cmp)> pq(cmp);
    std::unordered_set<std::string, KeyHash> seen;
    uint64_t seq = 0, configs = 0, steps = 0;
    Cfg c0{0, 0, m.init_value, std::vector<uint8_t>(sim.sh.n_classes, 0), 0, seq++};
    pq.push(c0);
    bool found = false;
    int max_rj = 0;
    while (!pq.empty() && !found && configs < max_configs) {
        std::vector<Cfg> batch;
        while (!pq.empty() && batch.size() < W) { batch.push_back(pq.top()); pq.pop(); }
        ++steps;
        for (const Cfg& c : batch) {
            std::vector<Cfg> kids;
            if (sim.expand(c, kids)) { found = true; break; }
            for (Cfg& k : kids)
                if (seen.insert(key_of(k)).second) {
                    ++configs;
                    max_rj = std::max(max_rj, k.rj);
                    k.seq = seq++;
                    if (bucketed) {
                        k.pb = k.rj >> shift;
                        int& mc = min_crashed[k.pb];
                        mc = std::min(mc, k.crashed_used);
                        k.ps = std::min(15, k.crashed_used - mc);
                    }
                    pq.push(std::move(k));
                }
        }
    }
    std::printf("{\"policy\": \"%s\", \"W\": %zu, \"found\": %s, \"exhausted\": %s, \"configs\": %llu, \"steps\": %llu, \"max_rank\": %d, \"ranks\": %zu}\n",
                policy.c_str(), W, found ? "true" : "false", (!found && pq.empty()) ? "true" : "false",
                (unsigned long long)configs, (unsigned long long)steps, max_rj, sim.sh.rets.size());
    return 0;
}
