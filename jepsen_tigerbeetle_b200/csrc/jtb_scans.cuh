// jtb_scans.cuh — the single-pass checkers on the hot path as HBM-streaming kernels.
//
//  K4 set_full:    jepsen.checker/set-full as called at src/tigerbeetle/workloads/set_full.clj:157
//                  ({:linearizable? true}); semantics SURVEY.md A.3.  Column-scan formulation: for
//                  element e (tracked from its last :add :invoke at index i_e) over :ok reads r with
//                  ok_idx[r] > i_e:
//                      last_present = max inv_idx[r] with e in r ; last_absent = max inv_idx[r] with e not in r
//                      known = min(first :add :ok after i_e, min ok_idx[r] with e in r)
//                  Stage A builds the read-major bit-matrix P[r][e] from the CSR id lists (binary search of each
//                  id in the shard's sorted element table, atomicOr); stage B scans columns (one thread per
//                  element, 32 elements share each 4 B word => coalesced broadcast loads); stage C classifies.
//  K5 bank_totals: src/tigerbeetle/tests/ledger.clj:127-192 (check-op precedence unexpected-key > nil-balance >
//                  wrong-total > negative-value; aggregation; err-badness :116-125).
#pragma once
#include <algorithm>
#include <chrono>
#include <cstdint>
#include <cstring>
#include <string>
#include <unordered_map>
#include <vector>

#include <cuda_runtime.h>

#include "../../include/jtb_check.h"

namespace jtb {

// =================================================================================================
// set-full
// =================================================================================================
struct SfRead {        // one :ok read (device)
    int32_t inv_idx, ok_idx;
    int64_t inv_time, ok_time;
    int64_t pl_off;    // into payload
    int32_t pl_len;
    int32_t shard;     // bit 31 set: :final? read
};
constexpr int32_t SF_FINAL_BIT = (int32_t)0x80000000;
struct SfShard {       // device
    int64_t elem_off;  // into element arrays
    int32_t n_elems;
    int32_t n_reads;
    int64_t read_off;  // into reads
    int64_t bits_off;  // into bit matrix (uint32 words); row stride = words_per_row
    int32_t words_per_row;
    int32_t pad;
};
struct SfElem {        // device, sorted by id within a shard
    int32_t id;
    int32_t add_inv_idx;   // last :add :invoke of this value
    int32_t add_ok_idx;    // first :add :ok after it, INT32_MAX if none
    int32_t order;         // creation order (output position within the shard)
    int64_t add_ok_time;
};
struct SfAcc {         // per element accumulators (device)
    unsigned long long last_present;  // (inv_idx+1) << 32 | read id   (0 = none)
    unsigned long long last_absent;
    unsigned long long known_read;    // min: ok_idx << 32 | read id    (~0 = none)
    int dup_max;
    int pad;
};

__global__ void sf_build_bits(const SfRead* __restrict__ reads, int64_t n_reads, const SfShard* __restrict__ shards,
                              const SfElem* __restrict__ elems, const int32_t* __restrict__ payload,
                              uint32_t* __restrict__ bits, int* __restrict__ read_dup_flag) {
    // one warp per read; lanes stride over the read's id list
    const int64_t r = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (r >= n_reads) return;
    const SfRead rd = reads[r];
    const SfShard sd = shards[rd.shard & ~SF_FINAL_BIT];
    const SfElem* el = elems + sd.elem_off;
    uint32_t* rowbits = bits + sd.bits_off + (r - sd.read_off) * (int64_t)sd.words_per_row;
    bool dup = false;
    for (int i = lane; i < rd.pl_len; i += 32) {
        const int32_t id = __ldg(payload + rd.pl_off + i);
        int lo = 0, hi = sd.n_elems - 1, pos = -1;
        while (lo <= hi) {
            const int mid = (lo + hi) >> 1;
            const int32_t v = __ldg(&el[mid].id);
            if (v == id) { pos = mid; break; }
            if (v < id) lo = mid + 1; else hi = mid - 1;
        }
        if (pos >= 0) {
            const uint32_t bit = 1u << (pos & 31);
            const uint32_t old = atomicOr(rowbits + (pos >> 5), bit);
            dup |= (old & bit) != 0;
        }
    }
    if (__any_sync(0xffffffffu, dup) && lane == 0) read_dup_flag[r] = 1;
}

// (read-all-invoked-adds) workloads/set_full.clj:51-75 on the same bit-matrix: a :final? :ok read is suspect
// when any tracked element (= any value ever :add-invoked in the sub-history) is absent from it.
// One warp per read; writes the number of missing elements (0 for non-final reads).
__global__ void sf_final_missing(const SfRead* __restrict__ reads, int64_t n_reads, const SfShard* __restrict__ shards,
                                 const uint32_t* __restrict__ bits, int* __restrict__ missing) {
    const int64_t r = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (r >= n_reads) return;
    const SfRead rd = reads[r];
    if (!(rd.shard & SF_FINAL_BIT)) { if (lane == 0) missing[r] = 0; return; }
    const SfShard sd = shards[rd.shard & ~SF_FINAL_BIT];
    const uint32_t* rowbits = bits + sd.bits_off + (r - sd.read_off) * (int64_t)sd.words_per_row;
    int zeros = 0;
    for (int w = lane; w < sd.words_per_row; w += 32) {
        const int valid_bits = min(32, sd.n_elems - w * 32);
        const uint32_t m = valid_bits == 32 ? 0xffffffffu : ((1u << valid_bits) - 1);
        zeros += __popc(~rowbits[w] & m);
    }
    for (int o = 16; o > 0; o >>= 1) zeros += __shfl_xor_sync(0xffffffffu, zeros, o);
    if (lane == 0) missing[r] = zeros;
}

// exact multiplicities for the (rare) reads that contain a repeated element
__global__ void sf_count_dups(const SfRead* __restrict__ reads, const int* __restrict__ flagged, int n_flagged,
                              const SfShard* __restrict__ shards, const SfElem* __restrict__ elems,
                              const int32_t* __restrict__ payload, SfAcc* __restrict__ acc) {
    const int r = flagged[blockIdx.x];
    if (blockIdx.x >= n_flagged) return;
    const SfRead rd = reads[r];
    const SfShard sd = shards[rd.shard & ~SF_FINAL_BIT];
    const SfElem* el = elems + sd.elem_off;
    for (int i = threadIdx.x; i < rd.pl_len; i += blockDim.x) {
        const int32_t id = payload[rd.pl_off + i];
        int cnt = 0;
        for (int j = 0; j < rd.pl_len; ++j) cnt += payload[rd.pl_off + j] == id;
        if (cnt > 1) {
            int lo = 0, hi = sd.n_elems - 1;
            while (lo <= hi) {
                const int mid = (lo + hi) >> 1;
                if (el[mid].id == id) { atomicMax(&acc[sd.elem_off + mid].dup_max, cnt); break; }
                if (el[mid].id < id) lo = mid + 1; else hi = mid - 1;
            }
        }
    }
}

constexpr int SF_RCHUNK = 2048;  // reads per block in the column scan

// grid: (element tiles of 256, read chunks, shards)
__global__ void __launch_bounds__(256) sf_column_scan(const SfRead* __restrict__ reads, const SfShard* __restrict__ shards,
                                                      const SfElem* __restrict__ elems, const uint32_t* __restrict__ bits,
                                                      SfAcc* __restrict__ acc) {
    const SfShard sd = shards[blockIdx.z];
    const int e = blockIdx.x * 256 + threadIdx.x;
    const int r0 = blockIdx.y * SF_RCHUNK;
    if (blockIdx.x * 256 >= sd.n_elems || r0 >= sd.n_reads) return;
    const int r1 = min(sd.n_reads, r0 + SF_RCHUNK);
    __shared__ int s_inv[256], s_ok[256];
    const bool live = e < sd.n_elems;
    const int add_inv = live ? elems[sd.elem_off + e].add_inv_idx : 0x7fffffff;
    unsigned long long lp = 0, la = 0, kn = ~0ull;
    const uint32_t* col = bits + sd.bits_off + (e >> 5);
    for (int rb = r0; rb < r1; rb += 256) {
        __syncthreads();
        if (rb + threadIdx.x < r1) {
            s_inv[threadIdx.x] = reads[sd.read_off + rb + threadIdx.x].inv_idx;
            s_ok[threadIdx.x] = reads[sd.read_off + rb + threadIdx.x].ok_idx;
        }
        __syncthreads();
        const int n = min(256, r1 - rb);
        if (live) {
#pragma unroll 4
            for (int k = 0; k < n; ++k) {
                const int r = rb + k;
                const uint32_t wv = __ldg(col + (int64_t)r * sd.words_per_row);
                if (s_ok[k] > add_inv) {
                    const bool present = (wv >> (e & 31)) & 1u;
                    const unsigned long long iv = ((unsigned long long)(uint32_t)(s_inv[k] + 1) << 32) | (uint32_t)r;
                    if (present) {
                        lp = max(lp, iv);
                        kn = min(kn, ((unsigned long long)(uint32_t)s_ok[k] << 32) | (uint32_t)r);
                    } else {
                        la = max(la, iv);
                    }
                }
            }
        }
    }
    if (live) {
        SfAcc* a = &acc[sd.elem_off + e];
        if (gridDim.y == 1) {
            a->last_present = lp; a->last_absent = la; a->known_read = kn;
        } else {
            if (lp) atomicMax(&a->last_present, lp);
            if (la) atomicMax(&a->last_absent, la);
            if (kn != ~0ull) atomicMin(&a->known_read, kn);
        }
    }
}

struct SfShardOut {   // device per-shard tallies
    int attempt, stable, lost, never_read, stale, duplicated;
    long long stable_lat_max, lost_lat_max;
};

__global__ void sf_classify(const SfRead* __restrict__ reads, const SfShard* __restrict__ shards,
                            const SfElem* __restrict__ elems, const SfAcc* __restrict__ acc, int64_t n_elems_total,
                            const int32_t* __restrict__ elem_shard, uint8_t* __restrict__ outcome,
                            long long* __restrict__ latency_ms, int* __restrict__ dup_count, int32_t* __restrict__ out_id,
                            SfShardOut* __restrict__ tallies) {
    const int64_t g = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (g >= n_elems_total) return;
    const int s = elem_shard[g];
    const SfShard sd = shards[s];
    const SfElem el = elems[g];
    const SfAcc a = acc[g];
    const bool has_lp = a.last_present != 0, has_la = a.last_absent != 0;
    const int lp_idx = has_lp ? (int)(a.last_present >> 32) - 1 : -1;
    const int la_idx = has_la ? (int)(a.last_absent >> 32) - 1 : -1;
    // known = earliest of (first add :ok after the tracking invoke, first observing read :ok)
    int known_idx = 0x7fffffff;
    long long known_time = 0;
    if (el.add_ok_idx != 0x7fffffff) { known_idx = el.add_ok_idx; known_time = el.add_ok_time; }
    if (a.known_read != ~0ull) {
        const int k = (int)(a.known_read >> 32);
        if (k < known_idx) { known_idx = k; known_time = reads[sd.read_off + (uint32_t)a.known_read].ok_time; }
    }
    const bool has_known = known_idx != 0x7fffffff;
    const bool stable = has_lp && la_idx < lp_idx;
    const bool lost = has_known && has_la && lp_idx < la_idx && known_idx < la_idx;
    int oc = JTB_SF_NEVER_READ;
    long long lat = 0;
    if (stable) {
        const long long stable_time = has_la ? reads[sd.read_off + (uint32_t)a.last_absent].inv_time + 1 : 0;
        const long long d = max(0ll, stable_time - known_time);
        lat = (long long)((double)d / 1e6);  // (long (util/nanos->ms d))
        oc = JTB_SF_STABLE;
        atomicAdd(&tallies[s].stable, 1);
        if (lat > 0) atomicAdd(&tallies[s].stale, 1);
        atomicMax(&tallies[s].stable_lat_max, lat);
    } else if (lost) {
        const long long lost_time = has_lp ? reads[sd.read_off + (uint32_t)a.last_present].inv_time + 1 : 0;
        const long long d = max(0ll, lost_time - known_time);
        lat = (long long)((double)d / 1e6);
        oc = JTB_SF_LOST;
        atomicAdd(&tallies[s].lost, 1);
        atomicMax(&tallies[s].lost_lat_max, lat);
    } else {
        atomicAdd(&tallies[s].never_read, 1);
    }
    if (a.dup_max > 1) atomicAdd(&tallies[s].duplicated, 1);
    // output in creation order within the shard
    const int64_t o = sd.elem_off + el.order;
    outcome[o] = (uint8_t)oc;
    latency_ms[o] = lat;
    dup_count[o] = a.dup_max > 1 ? a.dup_max : 0;
    out_id[o] = el.id;
}

#define SCK(call)                                                                  \
    do {                                                                           \
        cudaError_t e_ = (call);                                                   \
        if (e_ != cudaSuccess) {                                                   \
            err = std::string(#call) + ": " + cudaGetErrorString(e_);              \
            cleanup();                                                             \
            return -1;                                                             \
        }                                                                          \
    } while (0)

inline int run_set_full(cudaStream_t st, cudaEvent_t e0, cudaEvent_t e1, const jtb_history* h, int linearizable,
                        jtb_setfull_out* out, std::string& err) {
    const double t_start = std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
    const int n_shards = h->n_shards;
    // ---- host pass: tracked elements and :ok reads per shard (O(events)) -----------------------
    std::vector<SfShard> shards(n_shards);
    std::vector<SfElem> elems;
    std::vector<int32_t> elem_shard;
    std::vector<SfRead> reads;
    int64_t bits_words = 0;
    for (int s = 0; s < n_shards; ++s) {
        struct Tr { int32_t add_inv_idx, add_ok_idx, order; int64_t add_ok_time; };
        std::unordered_map<int32_t, Tr> tracked;
        std::unordered_map<int32_t, std::pair<int32_t, int64_t>> open_reads;  // process -> (inv idx, inv time)
        int n_created = 0;
        const int64_t read_off = (int64_t)reads.size();
        for (int64_t e = h->shard_off[s]; e < h->shard_off[s + 1]; ++e) {
            if (h->process[e] < 0) continue;
            const int type = h->type[e], f = h->f[e];
            if (f == JTB_F_ADD) {
                if (type == JTB_T_INVOKE) tracked[h->a[e]] = Tr{h->index[e], 0x7fffffff, n_created++, 0};
                else if (type == JTB_T_OK) {
                    auto it = tracked.find(h->a[e]);
                    if (it != tracked.end() && it->second.add_ok_idx == 0x7fffffff) {
                        it->second.add_ok_idx = h->index[e];
                        it->second.add_ok_time = h->time_ns[e];
                    }
                }
            } else if (f == JTB_F_READ) {
                const int32_t p = h->process[e];
                if (type == JTB_T_INVOKE) open_reads[p] = {h->index[e], h->time_ns[e]};
                else if (type == JTB_T_FAIL) open_reads.erase(p);
                else if (type == JTB_T_OK) {
                    auto it = open_reads.find(p);
                    if (it == open_reads.end()) { err = "malformed history: read :ok without invoke"; return -3; }
                    SfRead r;
                    r.inv_idx = it->second.first; r.inv_time = it->second.second;
                    r.ok_idx = h->index[e]; r.ok_time = h->time_ns[e];
                    r.pl_off = h->payload_off[e];
                    r.pl_len = std::max(0, (int)h->payload_len[e]);
                    r.shard = s | ((h->flags && (h->flags[e] & JTB_FLAG_FINAL)) ? SF_FINAL_BIT : 0);
                    reads.push_back(r);
                }
            }
        }
        // creation order is dense over the surviving elements
        std::vector<std::pair<int32_t, Tr>> v(tracked.begin(), tracked.end());
        std::sort(v.begin(), v.end(), [](auto& x, auto& y) { return x.second.order < y.second.order; });
        for (size_t i = 0; i < v.size(); ++i) v[i].second.order = (int32_t)i;
        std::sort(v.begin(), v.end(), [](auto& x, auto& y) { return x.first < y.first; });
        SfShard& sd = shards[s];
        sd.elem_off = (int64_t)elems.size();
        sd.n_elems = (int32_t)v.size();
        sd.read_off = read_off;
        sd.n_reads = (int32_t)(reads.size() - read_off);
        sd.words_per_row = (sd.n_elems + 31) / 32;
        sd.bits_off = bits_words;
        sd.pad = 0;
        bits_words += (int64_t)sd.words_per_row * sd.n_reads;
        for (auto& kv : v) {
            elems.push_back(SfElem{kv.first, kv.second.add_inv_idx, kv.second.add_ok_idx, kv.second.order, kv.second.add_ok_time});
            elem_shard.push_back(s);
        }
    }
    const int64_t n_elems = (int64_t)elems.size(), n_reads = (int64_t)reads.size();
    if (out->elem_capacity > 0 && out->elem_capacity < n_elems) { err = "elem_capacity too small"; return -4; }
    // ---- device buffers ---------------------------------------------------------------------------
    SfShard* d_shards = nullptr; SfElem* d_elems = nullptr; int32_t* d_elem_shard = nullptr; SfRead* d_reads = nullptr;
    int32_t* d_payload = nullptr; uint32_t* d_bits = nullptr; int* d_dupflag = nullptr; SfAcc* d_acc = nullptr;
    uint8_t* d_outcome = nullptr; long long* d_lat = nullptr; int* d_dup = nullptr; int32_t* d_id = nullptr;
    SfShardOut* d_tally = nullptr; int* d_flagged = nullptr; int* d_missing = nullptr;
    auto cleanup = [&]() {
        cudaFree(d_shards); cudaFree(d_elems); cudaFree(d_elem_shard); cudaFree(d_reads); cudaFree(d_payload);
        cudaFree(d_bits); cudaFree(d_dupflag); cudaFree(d_acc); cudaFree(d_outcome); cudaFree(d_lat); cudaFree(d_dup);
        cudaFree(d_id); cudaFree(d_tally); cudaFree(d_flagged); cudaFree(d_missing);
    };
    auto nz = [](size_t b) { return b ? b : (size_t)16; };
    SCK(cudaMalloc(&d_shards, nz(n_shards * sizeof(SfShard))));
    SCK(cudaMalloc(&d_elems, nz(n_elems * sizeof(SfElem))));
    SCK(cudaMalloc(&d_elem_shard, nz(n_elems * 4)));
    SCK(cudaMalloc(&d_reads, nz(n_reads * sizeof(SfRead))));
    SCK(cudaMalloc(&d_payload, nz((size_t)h->n_payload * 4)));
    SCK(cudaMalloc(&d_bits, nz((size_t)bits_words * 4)));
    SCK(cudaMalloc(&d_dupflag, nz(n_reads * 4)));
    SCK(cudaMalloc(&d_acc, nz(n_elems * sizeof(SfAcc))));
    SCK(cudaMalloc(&d_outcome, nz(n_elems)));
    SCK(cudaMalloc(&d_lat, nz(n_elems * 8)));
    SCK(cudaMalloc(&d_dup, nz(n_elems * 4)));
    SCK(cudaMalloc(&d_id, nz(n_elems * 4)));
    SCK(cudaMalloc(&d_tally, nz(n_shards * sizeof(SfShardOut))));
    SCK(cudaMalloc(&d_missing, nz(n_reads * 4)));
    SCK(cudaMemsetAsync(d_missing, 0, nz(n_reads * 4), st));
    SCK(cudaMemcpyAsync(d_shards, shards.data(), n_shards * sizeof(SfShard), cudaMemcpyHostToDevice, st));
    SCK(cudaMemcpyAsync(d_elems, elems.data(), n_elems * sizeof(SfElem), cudaMemcpyHostToDevice, st));
    SCK(cudaMemcpyAsync(d_elem_shard, elem_shard.data(), n_elems * 4, cudaMemcpyHostToDevice, st));
    SCK(cudaMemcpyAsync(d_reads, reads.data(), n_reads * sizeof(SfRead), cudaMemcpyHostToDevice, st));
    SCK(cudaMemcpyAsync(d_payload, h->payload, (size_t)h->n_payload * 4, cudaMemcpyHostToDevice, st));
    SCK(cudaEventRecord(e0, st));
    SCK(cudaMemsetAsync(d_bits, 0, nz((size_t)bits_words * 4), st));
    SCK(cudaMemsetAsync(d_dupflag, 0, nz(n_reads * 4), st));
    SCK(cudaMemsetAsync(d_tally, 0, nz(n_shards * sizeof(SfShardOut)), st));
    {   // accumulators: last_present = last_absent = 0, known_read = ~0, dup_max = 0
        std::vector<SfAcc> init((size_t)n_elems, SfAcc{0, 0, ~0ull, 0, 0});
        SCK(cudaMemcpyAsync(d_acc, init.data(), n_elems * sizeof(SfAcc), cudaMemcpyHostToDevice, st));
        SCK(cudaStreamSynchronize(st));  // `init` must outlive the copy
    }
    if (n_reads > 0 && n_elems > 0) {
        const int64_t threads = n_reads * 32;
        sf_build_bits<<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(d_reads, n_reads, d_shards, d_elems, d_payload, d_bits, d_dupflag);
        SCK(cudaGetLastError());
        int max_e = 0, max_r = 0;
        for (auto& sd : shards) { max_e = std::max(max_e, sd.n_elems); max_r = std::max(max_r, sd.n_reads); }
        dim3 grid((max_e + 255) / 256, (max_r + SF_RCHUNK - 1) / SF_RCHUNK, n_shards);
        sf_column_scan<<<grid, 256, 0, st>>>(d_reads, d_shards, d_elems, d_bits, d_acc);
        SCK(cudaGetLastError());
        sf_final_missing<<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(d_reads, n_reads, d_shards, d_bits, d_missing);
        SCK(cudaGetLastError());
        // duplicates (rare): exact multiplicities for flagged reads
        std::vector<int> flag((size_t)n_reads);
        SCK(cudaMemcpyAsync(flag.data(), d_dupflag, n_reads * 4, cudaMemcpyDeviceToHost, st));
        SCK(cudaStreamSynchronize(st));
        std::vector<int> flagged;
        for (int64_t r = 0; r < n_reads; ++r) if (flag[r]) flagged.push_back((int)r);
        if (!flagged.empty()) {
            SCK(cudaMalloc(&d_flagged, flagged.size() * 4));
            SCK(cudaMemcpyAsync(d_flagged, flagged.data(), flagged.size() * 4, cudaMemcpyHostToDevice, st));
            sf_count_dups<<<(unsigned)flagged.size(), 128, 0, st>>>(d_reads, d_flagged, (int)flagged.size(), d_shards, d_elems, d_payload, d_acc);
            SCK(cudaGetLastError());
        }
    }
    if (n_elems > 0) {
        sf_classify<<<(unsigned)((n_elems + 255) / 256), 256, 0, st>>>(d_reads, d_shards, d_elems, d_acc, n_elems, d_elem_shard,
                                                                      d_outcome, d_lat, d_dup, d_id, d_tally);
        SCK(cudaGetLastError());
    }
    SCK(cudaEventRecord(e1, st));
    std::vector<SfShardOut> tally(n_shards);
    SCK(cudaMemcpyAsync(tally.data(), d_tally, n_shards * sizeof(SfShardOut), cudaMemcpyDeviceToHost, st));
    if (out->elem_capacity > 0 && n_elems > 0) {
        SCK(cudaMemcpyAsync(out->elem_outcome, d_outcome, n_elems, cudaMemcpyDeviceToHost, st));
        SCK(cudaMemcpyAsync(out->elem_latency_ms, d_lat, n_elems * 8, cudaMemcpyDeviceToHost, st));
        SCK(cudaMemcpyAsync(out->elem_dup_count, d_dup, n_elems * 4, cudaMemcpyDeviceToHost, st));
        SCK(cudaMemcpyAsync(out->elem_id, d_id, n_elems * 4, cudaMemcpyDeviceToHost, st));
    }
    SCK(cudaStreamSynchronize(st));
    float ms = 0;
    SCK(cudaEventElapsedTime(&ms, e0, e1));
    // read-all-invoked-adds: suspects = final reads with missing elements; ids enumerated from their bit rows
    std::vector<int> h_missing((size_t)n_reads, 0);
    if (n_reads > 0) SCK(cudaMemcpy(h_missing.data(), d_missing, n_reads * 4, cudaMemcpyDeviceToHost));
    std::vector<int> suspect_per_shard(n_shards, 0);
    out->n_suspect = 0;
    out->raia_valid = JTB_VALID;
    {
        int64_t cursor = 0;
        if (out->suspect_capacity > 0) out->suspect_missing_off[0] = 0;
        std::vector<uint32_t> rowbits;
        for (int64_t r = 0; r < n_reads; ++r) {
            if (!h_missing[r]) continue;
            const int s = reads[r].shard & ~SF_FINAL_BIT;
            suspect_per_shard[s]++;
            out->raia_valid = JTB_INVALID;
            if (out->suspect_capacity > 0) {
                if (out->n_suspect >= out->suspect_capacity || cursor + h_missing[r] > out->missing_capacity) {
                    err = "suspect/missing capacity too small";
                    cleanup();
                    return -4;
                }
                const SfShard& sd = shards[s];
                rowbits.resize(sd.words_per_row);
                SCK(cudaMemcpy(rowbits.data(), d_bits + sd.bits_off + (r - sd.read_off) * (int64_t)sd.words_per_row,
                               (size_t)sd.words_per_row * 4, cudaMemcpyDeviceToHost));
                out->suspect_shard[out->n_suspect] = s;
                out->suspect_index[out->n_suspect] = reads[r].ok_idx;
                for (int e = 0; e < sd.n_elems; ++e)  // elems are sorted by id inside a shard
                    if (!((rowbits[e >> 5] >> (e & 31)) & 1u)) out->missing_ids[cursor++] = elems[sd.elem_off + e].id;
                out->suspect_missing_off[out->n_suspect + 1] = cursor;
            }
            out->n_suspect++;
        }
    }
    out->valid = JTB_VALID;
    out->n_failures = 0;
    if (out->elem_capacity > 0) out->elem_off[0] = 0;
    for (int s = 0; s < n_shards; ++s) {
        jtb_setfull_shard& r = out->shards[s];
        std::memset(&r, 0, sizeof r);
        r.suspect_final_reads = suspect_per_shard[s];
        r.attempt_count = shards[s].n_elems;
        r.stable_count = tally[s].stable; r.lost_count = tally[s].lost; r.never_read_count = tally[s].never_read;
        r.stale_count = tally[s].stale; r.duplicated_count = tally[s].duplicated;
        r.stable_latency_max_ms = tally[s].stable_lat_max; r.lost_latency_max_ms = tally[s].lost_lat_max;
        int valid;
        if (r.lost_count > 0) valid = JTB_INVALID;
        else if (r.stable_count == 0) valid = JTB_UNKNOWN;
        else if (linearizable && r.stale_count > 0) valid = JTB_INVALID;
        else valid = JTB_VALID;
        if (r.duplicated_count > 0) valid = JTB_INVALID;
        r.valid = valid;
        out->valid = std::max(out->valid, valid);
        out->n_failures += valid != JTB_VALID;
        if (out->elem_capacity > 0) out->elem_off[s + 1] = shards[s].elem_off + shards[s].n_elems;
    }
    cleanup();
    out->seconds_kernel = ms * 1e-3;
    out->seconds_total = std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count() - t_start;
    return 0;
}

// =================================================================================================
// bank totals
// =================================================================================================
struct BkRead { int64_t pl_off; int32_t pl_len; int32_t index; };
struct BkAgg {
    unsigned long long count[5];
    int first_idx[5], last_idx[5];
    unsigned long long worst_key[5];   // orderable badness
    int worst_idx[5];
    long long lowest, highest;
    int lowest_idx, highest_idx;
    int first_error_idx;
    int pad;
};

__device__ __forceinline__ unsigned long long orderable(double d) {
    unsigned long long u = (unsigned long long)__double_as_longlong(d);
    return (u >> 63) ? ~u : (u | (1ull << 63));
}

// pass 0: classify + value reductions; pass 1: arg-index resolution (first read attaining the extreme)
__global__ void bk_scan(const BkRead* __restrict__ reads, int64_t n_reads, const int32_t* __restrict__ payload,
                        int n_accounts, int4 acct_lo, int4 acct_hi, long long total_amount, int neg_ok, int pass,
                        uint8_t* __restrict__ type_out, BkAgg* __restrict__ agg) {
    const int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (r >= n_reads) return;
    const BkRead rd = reads[r];
    const int ids[8] = {acct_lo.x, acct_lo.y, acct_lo.z, acct_lo.w, acct_hi.x, acct_hi.y, acct_hi.z, acct_hi.w};
    int n_unexpected = 0, n_nil = 0;
    long long total = 0, neg_sum = 0;
    bool any_neg = false;
    for (int i = 0; i + 1 < rd.pl_len; i += 2) {
        const int32_t id = __ldg(payload + rd.pl_off + i), bal = __ldg(payload + rd.pl_off + i + 1);
        bool known = false;
#pragma unroll
        for (int a = 0; a < 8; ++a) known |= a < n_accounts && ids[a] == id;
        n_unexpected += !known;
        if (bal == JTB_NIL) n_nil++;
        else { total += bal; if (bal < 0) { any_neg = true; neg_sum += bal; } }
    }
    int type = JTB_BANK_OK;
    if (n_unexpected) type = JTB_BANK_UNEXPECTED_KEY;
    else if (n_nil) type = JTB_BANK_NIL_BALANCE;
    else if (total != total_amount) type = JTB_BANK_WRONG_TOTAL;
    else if (!neg_ok && any_neg) type = JTB_BANK_NEGATIVE_VALUE;
    if (type == JTB_BANK_OK) { if (pass == 0) type_out[r] = 0; return; }
    double bad = 0;
    if (type == JTB_BANK_UNEXPECTED_KEY) bad = n_unexpected;
    else if (type == JTB_BANK_NIL_BALANCE) bad = n_nil;
    else if (type == JTB_BANK_WRONG_TOTAL)
        bad = total_amount == 0 ? fabs((double)(total - total_amount))
                                : fabs((double)(float)((double)(total - total_amount) / (double)total_amount));
    else bad = -(double)neg_sum;
    const unsigned long long bk = orderable(bad);
    if (pass == 0) {
        type_out[r] = (uint8_t)type;
        atomicAdd(&agg->count[type], 1ull);
        atomicMin(&agg->first_idx[type], rd.index);
        atomicMax(&agg->last_idx[type], rd.index);
        atomicMax(&agg->worst_key[type], bk);
        atomicMin(&agg->first_error_idx, rd.index);
        if (type == JTB_BANK_WRONG_TOTAL) { atomicMin(&agg->lowest, total); atomicMax(&agg->highest, total); }
    } else {
        if (agg->worst_key[type] == bk) atomicMin(&agg->worst_idx[type], rd.index);
        if (type == JTB_BANK_WRONG_TOTAL) {
            if (agg->lowest == total) atomicMin(&agg->lowest_idx, rd.index);
            if (agg->highest == total) atomicMin(&agg->highest_idx, rd.index);
        }
    }
}

inline int run_bank_totals(cudaStream_t st, cudaEvent_t e0, cudaEvent_t e1, const jtb_history* h, const jtb_model* m,
                           int64_t total_amount, jtb_bank_result* out, std::string& err) {
    const double t_start = std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
    std::vector<BkRead> reads;
    for (int64_t e = 0; e < h->n_events; ++e) {
        if (h->process[e] < 0 || h->type[e] != JTB_T_OK || h->f[e] != JTB_F_READ) continue;
        reads.push_back(BkRead{h->payload_off[e], std::max(0, (int)h->payload_len[e]), h->index[e]});
    }
    const int64_t n = (int64_t)reads.size();
    BkRead* d_reads = nullptr; int32_t* d_payload = nullptr; uint8_t* d_type = nullptr; BkAgg* d_agg = nullptr;
    auto cleanup = [&]() { cudaFree(d_reads); cudaFree(d_payload); cudaFree(d_type); cudaFree(d_agg); };
    auto nz = [](size_t b) { return b ? b : (size_t)16; };
    SCK(cudaMalloc(&d_reads, nz(n * sizeof(BkRead))));
    SCK(cudaMalloc(&d_payload, nz((size_t)h->n_payload * 4)));
    SCK(cudaMalloc(&d_type, nz(n)));
    SCK(cudaMalloc(&d_agg, sizeof(BkAgg)));
    BkAgg init;
    std::memset(&init, 0, sizeof init);
    for (int t = 0; t < 5; ++t) { init.first_idx[t] = 0x7fffffff; init.last_idx[t] = -1; init.worst_idx[t] = 0x7fffffff; }
    init.lowest = INT64_MAX; init.highest = INT64_MIN;
    init.lowest_idx = init.highest_idx = init.first_error_idx = 0x7fffffff;
    SCK(cudaMemcpyAsync(d_reads, reads.data(), n * sizeof(BkRead), cudaMemcpyHostToDevice, st));
    SCK(cudaMemcpyAsync(d_payload, h->payload, (size_t)h->n_payload * 4, cudaMemcpyHostToDevice, st));
    SCK(cudaMemcpyAsync(d_agg, &init, sizeof init, cudaMemcpyHostToDevice, st));
    int ids[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int i = 0; i < m->n_accounts && i < 8; ++i) ids[i] = m->account_ids[i];
    const int4 lo = make_int4(ids[0], ids[1], ids[2], ids[3]), hi = make_int4(ids[4], ids[5], ids[6], ids[7]);
    SCK(cudaEventRecord(e0, st));
    if (n > 0) {
        for (int pass = 0; pass < 2; ++pass) {
            bk_scan<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(d_reads, n, d_payload, m->n_accounts, lo, hi, total_amount,
                                                                 m->negative_balances_ok, pass, d_type, d_agg);
            SCK(cudaGetLastError());
        }
    }
    SCK(cudaEventRecord(e1, st));
    BkAgg agg;
    SCK(cudaMemcpyAsync(&agg, d_agg, sizeof agg, cudaMemcpyDeviceToHost, st));
    std::vector<uint8_t> types((size_t)n);
    SCK(cudaMemcpyAsync(types.data(), d_type, n, cudaMemcpyDeviceToHost, st));
    SCK(cudaStreamSynchronize(st));
    float ms = 0;
    SCK(cudaEventElapsedTime(&ms, e0, e1));
    std::memset(out, 0, sizeof *out);
    out->read_count = n;
    out->first_error_index = -1;
    out->lowest_index = out->highest_index = -1;
    for (int t = 0; t < 5; ++t) {
        out->count_by_type[t] = (int64_t)agg.count[t];
        out->error_count += (t > 0) ? (int64_t)agg.count[t] : 0;
        out->first_index_by_type[t] = agg.count[t] ? agg.first_idx[t] : -1;
        out->last_index_by_type[t] = agg.count[t] ? agg.last_idx[t] : -1;
        out->worst_index_by_type[t] = agg.count[t] ? agg.worst_idx[t] : -1;
    }
    if (out->error_count) {
        out->first_error_index = agg.first_error_idx;
        for (int64_t r = 0; r < n; ++r)
            if (reads[r].index == agg.first_error_idx) { out->first_error_type = types[r]; break; }
    }
    if (agg.count[JTB_BANK_WRONG_TOTAL]) {
        out->lowest_total = agg.lowest; out->highest_total = agg.highest;
        out->lowest_index = agg.lowest_idx; out->highest_index = agg.highest_idx;
    }
    out->valid = out->error_count ? JTB_INVALID : JTB_VALID;
    // reference quirk (tests/ledger.clj:122-123 with the default :total-amount 0, :356): err-badness divides by zero as
    // soon as util/max-by has two :wrong-total errors to compare -> the checker throws -> check-safe: :unknown
    if (total_amount == 0 && agg.count[JTB_BANK_WRONG_TOTAL] >= 2) {
        out->reference_throws = 1;
        out->valid = JTB_UNKNOWN;
    }
    cleanup();
    out->seconds_kernel = ms * 1e-3;
    out->seconds_total = std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count() - t_start;
    return 0;
}

}  // namespace jtb
