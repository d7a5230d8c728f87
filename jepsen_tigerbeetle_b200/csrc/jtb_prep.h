// jtb_prep.h — host-side preparation of a flattened history for the device search.
//
// Product code (libjtb_check.so).  Independent of oracle/ (which is test infrastructure).
// Follows knossos.history/{complete, without-failures, pair-index}:
//   client ops only; invoke paired with completion by :process; :fail pairs removed; :info ops stay
//   open forever; crashed reads dropped.  Reference consumers of the same pairing:
//   src/tigerbeetle/tests/ledger.clj:206 (history/unmatched-invokes), checker/perf.clj:617,623.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "../../include/jtb_check.h"

namespace jtb {

// One op as the device sees it (16 B, loaded as int4).
//   x = f | flags<<8   (flag bit 8: impossible — can never be linearized)
//   register/cas: y = value / cas-old, z = cas-new, w = invocation position
//   bank transfer: y = amount, z = debit slot, w = credit slot (invocation position: row cell word 4)
//   bank read:     y = care mask over account slots (the balances follow inline in the frontier row), w = inv. pos.
//   set add:       y = element value, w = invocation position
//   set read:      (need, care) masks follow inline in the frontier row, per rank; w = invocation position
struct OpRec {
    int32_t x, y, z, w;
};
constexpr int OP_IMPOSSIBLE = 1 << 8;
// bank reads that cover every account: OpRec.z holds bank_hash() of the expected balances, so a candidate read is
// rejected with one 32-bit compare (the hash is linear in the balances: equal balances => equal hash)
constexpr int OP_HASHED = 1 << 9;
#if defined(__CUDACC__)
__host__ __device__
#endif
constexpr uint32_t bank_hash_c(int i) {
    return i == 0 ? 0x9E3779B1u : i == 1 ? 0x85EBCA77u : i == 2 ? 0xC2B2AE3Du : i == 3 ? 0x27D4EB2Fu :
           i == 4 ? 0x165667B1u : i == 5 ? 0xD3A2646Du : i == 6 ? 0xFD7046C5u : 0xB55A4F09u;
}

// Crashed-op equivalence class (same f and value): members are linearized in invocation order only,
// so a config records just how many of the class it has consumed (a count field inside the key).
struct ClassRec {
    OpRec op;
    int32_t first;   // offset of the class' members in cls_inv_pos
    int32_t n;       // number of members
    int32_t word;    // key word holding the count
    int32_t shift_width;  // shift | width << 8
};

// Row of the frontier table: everything a warp needs to expand a config whose first un-linearized
// return is global rank gj, in ONE contiguous read.  int32 words:
//   [0, 8)     slots (u8) of the next 32 returns gj+1 .. gj+32 (0xFF beyond the shard end)
//   [8]        position of the return event (for crashed-op eligibility)
//   [9]        shard id
//   [10]       global rank one past the shard's last return (success when reached)
//   [11]       first class record of the shard
//   [12]       number of classes of the shard
//   [13]       slot of this rank's own op
//   [14..15]   u64 mask of the slots that hold a linearizable op at this return (empty / impossible ops excluded)
//   [16..17]   u64 mask of those slots whose op is a READ (never changes the model state)
//   [18..19]   u64 mask of the READ slots that the summary word below decides (bank: reads covering every account,
//              summary = hash of the expected balances, equal hash still verified against the cell; register /
//              cas-register: every read, summary = the expected value, exact)
//   [20 + t*SW, 20 + (t+1)*SW)   the op occupying open-op slot t at that return event, INLINE:
//        words 0..3  OpRec (x = -1: slot empty)
//        bank  (SW = 12): words 4..11 = the 8 balances a read expects (by account slot; op.y = care mask);
//                         word 4 of a TRANSFER = its invocation position
//        set   (SW = 8):  words 4..7  = (need, care) u64 pair of a read AT THIS RANK:
//                         consistent <=> (key word1 & care) == need
//   [20 + S_pad*SW, 20 + S_pad*SW + S_pad)   one SUMMARY word per slot (see [18..19]): a thread decides every
//        candidate read of a configuration from this one contiguous array instead of one cell load per read
constexpr int ROW_EXTRA = 20;
inline int slot_words(int model) { return model == JTB_MODEL_BANK ? 12 : model == JTB_MODEL_SET ? 8 : 4; }
constexpr int OP_EMPTY = -1;

struct Prepared {
    int S_pad = 32;          // 32 or 64
    int key_words = 2;       // 64-bit words per key: 2, 4 or 8
    int model = 0;
    int64_t n_ranks = 0;     // total completed ops over all searchable shards
    int row_words = 0;                // ROW_EXTRA + S_pad * slot_words(model) + S_pad
    int sum_off = 0;                  // ROW_EXTRA + S_pad * slot_words(model): offset of the summary words in a row
    std::vector<int32_t> rows;        // n_ranks * row_words
    std::vector<OpRec> ops;           // global op table (host side only: copied inline into rows)
    std::vector<int32_t> read_bal;    // bank: 8 per read (host side only)
    std::vector<ClassRec> classes;
    std::vector<int32_t> cls_inv_pos;
    // per shard (host side)
    std::vector<int64_t> rank_base;   // [n_shards+1]
    std::vector<int32_t> ret_index;   // [n_ranks] :index of each return completion
    std::vector<int32_t> shard_cause; // JTB_CAUSE_* decided on the host (too wide), else 0
    std::vector<int32_t> max_classes; // classes per shard
    int max_nc = 0;                   // max classes in any shard
    double mean_open = 0;             // mean number of open (linearizable) ops per frontier row: how wide the search gets
    std::string error;
};

// Returns false (and sets out.error) on malformed histories.
bool prepare(const jtb_history* h, const jtb_model* m, Prepared& out);

}  // namespace jtb
