/*
 * jtb_check.h — C ABI of the B200-native history checker (libjtb_check.so).
 *
 * This is the drop-in boundary for the ONE hot path of nurturenature/jepsen-tigerbeetle that this
 * repo accelerates: the history checkers that sit behind the Clojure protocol
 * `jepsen.checker/Checker` (`(check [this test history opts]) -> {:valid? ...}`), composed by the
 * reference at
 *     src/tigerbeetle/workloads/set_full.clj:155-158   (independent/checker ∘ compose{set-full, read-all-invoked-adds})
 *     src/tigerbeetle/tests/ledger.clj:363-367         (compose{:SI checker, ...})
 *     src/tigerbeetle/core.clj:139-146                 (top-level compose)
 *
 * A JVM host (JNI) or any FFI binds exactly these entry points; all arguments are plain pointers and
 * sizes owned by the caller for the duration of the call.  Nothing here is a torch type.
 *
 * Verdict coding follows jepsen.checker/merge-valid (false dominates :unknown dominates true):
 *     JTB_VALID (0) < JTB_UNKNOWN (1) < JTB_INVALID (2)   so that merging == max.
 */
#ifndef JTB_CHECK_H
#define JTB_CHECK_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define JTB_ABI_VERSION 2

/* ---- verdict lattice (jepsen.checker/merge-valid) ------------------------------------------- */
#define JTB_VALID   0
#define JTB_UNKNOWN 1
#define JTB_INVALID 2

/* ---- op :type (knossos.op/{invoke?,ok?,fail?,info?}) ----------------------------------------- */
#define JTB_T_INVOKE 0
#define JTB_T_OK     1
#define JTB_T_FAIL   2
#define JTB_T_INFO   3

/* ---- op :f opcodes ----------------------------------------------------------------------------
 * register / cas-register (knossos.model):  READ a=value|JTB_NIL ; WRITE a=value ; CAS a=old b=new
 * set (knossos.model/set, jepsen.checker/set-full; set_full.clj:29-31,128-134):
 *                                           ADD a=element ; READ payload=element ids
 * bank (tests/ledger.clj:89-114 after ledger->bank):
 *                                           TRANSFER a=amount b=debit-acct c=credit-acct ;
 *                                           READ payload=(account id, balance) pairs
 */
#define JTB_F_READ     0
#define JTB_F_WRITE    1
#define JTB_F_CAS      2
#define JTB_F_ADD      3
#define JTB_F_TRANSFER 4

#define JTB_NIL INT32_MIN /* Clojure nil in an int32 field (a nil register read matches any state) */

/* ---- op flags --------------------------------------------------------------------------------- */
#define JTB_FLAG_FINAL 1u /* :final? true  (set_full.clj:45, tests/ledger.clj:78,84) */

/* ---- models for the linearizability search ---------------------------------------------------- */
#define JTB_MODEL_REGISTER     0 /* knossos.model/register      */
#define JTB_MODEL_CAS_REGISTER 1 /* knossos.model/cas-register  */
#define JTB_MODEL_SET          2 /* knossos.model/set (grow-only) */
#define JTB_MODEL_BANK         3 /* bank-transfer model implied by tests/ledger.clj:89-152 */

#define JTB_MAX_ACCOUNTS 8 /* core.clj:208-210 default (vec (range 1 9)) */

/*
 * Flattened history, struct-of-arrays, little-endian, events in history (:index) order.
 * Independent keys (jepsen.independent tuples, set_full.clj:31,44,116,134) are a CSR partition:
 * shard s owns events [shard_off[s], shard_off[s+1]); inside a shard events keep history order.
 * A history without independent keys has n_shards = 1, shard_off = {0, n_events}.
 * Events of non-client processes (:nemesis) carry process < 0 and are ignored by every checker
 * (tests/ledger.clj:94,204,228,262).
 */
typedef struct jtb_history {
    int64_t n_events;
    const uint8_t*  type;        /* [n_events] JTB_T_*                                              */
    const uint8_t*  f;           /* [n_events] JTB_F_*                                              */
    const uint8_t*  flags;       /* [n_events] JTB_FLAG_* (may be NULL = all zero)                  */
    const int32_t*  process;     /* [n_events] :process, <0 = not a client                          */
    const int32_t*  index;       /* [n_events] original :index (reported back as witness)           */
    const int64_t*  time_ns;     /* [n_events] :time                                                */
    const int32_t*  a;           /* [n_events] see opcodes                                          */
    const int32_t*  b;           /* [n_events]                                                      */
    const int32_t*  c;           /* [n_events]                                                      */
    const int64_t*  payload_off; /* [n_events] offset into payload (in int32 units)                 */
    const int32_t*  payload_len; /* [n_events] number of int32 in this event's payload, -1 = nil    */
    const int32_t*  payload;     /* [n_payload]                                                     */
    int64_t n_payload;
    int32_t n_shards;
    const int64_t*  shard_off;   /* [n_shards+1]                                                    */
    const int64_t*  key_ids;     /* [n_shards] the independent key of each shard (may be NULL)      */
} jtb_history;

typedef struct jtb_model {
    int32_t kind;                         /* JTB_MODEL_*                                            */
    int32_t init_value;                   /* register / cas-register initial value (JTB_NIL = nil)  */
    int32_t n_accounts;                   /* bank: (count (:accounts test)) <= JTB_MAX_ACCOUNTS     */
    int32_t account_ids[JTB_MAX_ACCOUNTS];/* bank: (:accounts test), core.clj:208-210               */
    int32_t init_balance[JTB_MAX_ACCOUNTS];/* bank: starting balances (db.clj:118-127 => zeros)     */
    int32_t negative_balances_ok;         /* bank: (:negative-balances? test), core.clj:217-219     */
} jtb_model;

/* By default the search linearizes a consistent candidate READ immediately and exclusively ("eager
 * reads": a read never changes the model state, so verdict and witness are unchanged while the number
 * of configurations drops by an order of magnitude).  Knossos does not do this; set this flag to visit
 * exactly the configurations Knossos' WGL would. */
#define JTB_OPT_NO_EAGER_READS 1
/* For histories with crashed (:info) ops a few depth-first "scout" warps walk the same configuration space in
 * knossos.wgl's order (and three other orders) beside the exhaustive search, each with a private visited table;
 * a scout can only ever report VALID (it found a linearization), so verdicts, witnesses and exhaustive
 * configuration counts are unaffected.  Set this flag to run the exhaustive search alone. */
#define JTB_OPT_NO_SCOUTS 2
/* Engine choice for jtb_check_linearizable.  Default (neither bit): the level-synchronous engine (csrc/jtb_level.cuh:
 * breadth-first by depth, visited set local to a level, bounded memory) for histories without crashed (:info) ops
 * that are searched in the Knossos-exact space or are wide (nearly every client always has an op in flight), the
 * work-list engine (csrc/jtb_wgl.cuh: depth-first locally, persistent visited table, scouts) otherwise.
 * Verdict, witness and exhaustive configuration counts are identical; the bits force one engine. */
#define JTB_OPT_ENGINE_LEVEL    4
#define JTB_OPT_ENGINE_WORKLIST 8
/* A single-key history with crashed (:info) ops first gets a budgeted run of the work list (16 M configurations) and,
 * if that leaves it open, a BEAM (the level engine expanding only the best ~256, then ~2k, then ~16k configurations of
 * every level: fewest crashed ops consumed, furthest frontier): it finds the linearization of a valid history in tens of
 * milliseconds where an exhaustive search drowns.  A beam can only ever report VALID; what it does not decide goes to
 * the exhaustive search.  Set this flag to skip both. */
#define JTB_OPT_NO_BEAM         16

/* Options for a context.  Zero-initialise, then set what you need. */
typedef struct jtb_opts {
    int32_t  device;            /* CUDA device ordinal for this context                              */
    int32_t  flags;             /* JTB_OPT_* bits                                                     */
    uint64_t table_bytes;       /* visited-config table size in HBM (0 = default 4 GiB)              */
    uint64_t max_configs;       /* search budget: stop with JTB_UNKNOWN after this many (0 = table)   */
    uint32_t time_budget_ms;    /* 0 = unlimited                                                      */
    uint32_t search_ctas;       /* 0 = one persistent CTA per SM × resident CTAs                      */
} jtb_opts;

/* Per-shard output of the linearizability search (knossos analysis map). */
typedef struct jtb_lin_shard {
    int32_t  valid;             /* JTB_VALID / JTB_UNKNOWN / JTB_INVALID                              */
    int32_t  witness_index;     /* :index of the :ok completion that cannot be linearized (:op), -1   */
    int32_t  previous_ok_index; /* :index of the last :ok completion before it (:previous-ok), -1     */
    int32_t  cause;             /* JTB_CAUSE_* when valid == JTB_UNKNOWN                              */
    uint64_t configs_explored;  /* distinct (linearized-set, model-state) configs inserted            */
    uint64_t probes;            /* visited-table probes (hits + misses)                               */
} jtb_lin_shard;

#define JTB_CAUSE_NONE          0
#define JTB_CAUSE_TABLE_FULL    1 /* visited table exhausted (knossos: out of memory -> :unknown)    */
#define JTB_CAUSE_BUDGET        2 /* max_configs / time budget reached                                */
#define JTB_CAUSE_TOO_WIDE      3 /* > 64 concurrently open completed ops, or key does not fit        */

typedef struct jtb_lin_result {
    int32_t  valid;             /* merge-valid over shards                                            */
    int32_t  n_failures;        /* number of shards whose verdict is not JTB_VALID (:failures)        */
    uint64_t configs_explored;  /* sum over shards                                                    */
    uint64_t probes;            /* sum over shards                                                    */
    uint64_t hbm_bytes_algorithmic; /* key_bytes*(probes + inserts), SURVEY §8(d)                     */
    uint32_t key_bytes;         /* bytes per visited-table slot used by this call (16/32/64)          */
    uint32_t reserved0;
    double   seconds_kernel;    /* device time of the search kernels (CUDA events)                    */
    double   seconds_total;     /* host wall time of the call incl. flatten-prep, H2D, D2H            */
} jtb_lin_result;

/* Per-shard output of jepsen.checker/set-full. Element lists are returned through
 * caller-provided buffers in jtb_setfull_out. */
typedef struct jtb_setfull_shard {
    int32_t valid;
    int32_t attempt_count, stable_count, lost_count, never_read_count, stale_count, duplicated_count;
    int32_t suspect_final_reads;   /* read-all-invoked-adds: :final? :ok reads missing an invoked add   */
    int64_t stable_latency_max_ms; /* max stable-latency (0 if none)  */
    int64_t lost_latency_max_ms;   /* max lost-latency (0 if none)    */
} jtb_setfull_shard;

/* Per-element classification codes written to jtb_setfull_out.elem_outcome */
#define JTB_SF_NEVER_READ 0
#define JTB_SF_STABLE     1
#define JTB_SF_LOST       2

typedef struct jtb_setfull_out {
    jtb_setfull_shard* shards;     /* [n_shards] required                                             */
    /* optional per-element detail, CSR by shard over tracked elements in first-add-invoke order:     */
    int64_t  elem_capacity;        /* capacity of the arrays below (0 = not wanted)                   */
    int64_t* elem_off;             /* [n_shards+1]                                                    */
    int32_t* elem_id;              /* element value                                                   */
    uint8_t* elem_outcome;         /* JTB_SF_*                                                        */
    int64_t* elem_latency_ms;      /* stable-latency / lost-latency in ms (0 for never-read)          */
    int32_t* elem_dup_count;       /* max multiplicity seen in one read if > 1, else 0                */
    int32_t  valid;                /* merge-valid over shards                                          */
    int32_t  n_failures;
    double   seconds_kernel;
    double   seconds_total;
    /* (read-all-invoked-adds), workloads/set_full.clj:51-75, evaluated in the same pass: every
     * :final? :ok read must contain every :add value ever invoked in its sub-history.  Optional detail
     * (CSR: suspect read -> missing element ids), caller-allocated:                                   */
    int64_t  suspect_capacity;     /* capacity of suspect_shard / suspect_index (0 = not wanted)       */
    int32_t* suspect_shard;
    int32_t* suspect_index;        /* :index of the suspect final read                                 */
    int64_t* suspect_missing_off;  /* [suspect_capacity + 1]                                           */
    int64_t  missing_capacity;
    int32_t* missing_ids;
    int64_t  n_suspect;            /* out: total suspect final reads                                   */
    int32_t  raia_valid;           /* out: JTB_VALID or JTB_INVALID                                    */
    int32_t  reserved1;
} jtb_setfull_out;

/* bank SI checker (tests/ledger.clj:127-192) error classes, in `cond` precedence order */
#define JTB_BANK_OK             0
#define JTB_BANK_UNEXPECTED_KEY 1
#define JTB_BANK_NIL_BALANCE    2
#define JTB_BANK_WRONG_TOTAL    3
#define JTB_BANK_NEGATIVE_VALUE 4

typedef struct jtb_bank_result {
    int32_t valid;                 /* JTB_UNKNOWN when reference_throws (what check-safe makes of the exception) */
    int32_t reference_throws;      /* 1: the reference checker THROWS on this history — err-badness divides by
                                      (:total-amount test) (tests/ledger.clj:122-123), the default is 0
                                      (tests/ledger.clj:356), and util/max-by calls it as soon as one error type has
                                      >= 2 :wrong-total errors; jepsen's check-safe turns that into
                                      {:valid? :unknown}.  All counts / firsts / lasts below are still filled;
                                      :worst then ranks :wrong-total errors by |total - total-amount|.            */
    int64_t read_count;            /* :read-count  */
    int64_t error_count;           /* :error-count */
    int32_t first_error_index;     /* :index of (:op :first-error), -1                                 */
    int32_t first_error_type;      /* JTB_BANK_*                                                        */
    int64_t count_by_type[5];      /* (:count (errors type))                                            */
    int32_t first_index_by_type[5];/* :index of :first                                                  */
    int32_t last_index_by_type[5]; /* :index of :last                                                   */
    int32_t worst_index_by_type[5];/* :index of :worst (err-badness, tests/ledger.clj:116-125)          */
    int64_t lowest_total, highest_total;       /* :wrong-total :lowest / :highest totals                */
    int32_t lowest_index, highest_index;
    double  seconds_kernel;
    double  seconds_total;
} jtb_bank_result;

typedef struct jtb_ctx jtb_ctx;

/* ---- lifecycle -------------------------------------------------------------------------------- */
int         jtb_abi_version(void);
/* sizeof of the ABI structs as this library was compiled, for binding self-checks:
 * 0 jtb_history, 1 jtb_model, 2 jtb_opts, 3 jtb_lin_shard, 4 jtb_lin_result, 5 jtb_setfull_shard,
 * 6 jtb_setfull_out, 7 jtb_bank_result, 8 jtb_final_config; -1 otherwise */
long        jtb_struct_size(int which);
int         jtb_device_count(void);                 /* number of CUDA devices, <0 on error          */
jtb_ctx*    jtb_create(const jtb_opts* opts);       /* NULL on failure (no CUDA device etc.)        */
void        jtb_destroy(jtb_ctx* ctx);
const char* jtb_last_error(const jtb_ctx* ctx);     /* valid until the next call on ctx             */

/* ---- hot path A9: jepsen.checker/linearizable -> knossos analysis ---------------------------- *
 * Replaces (checker/linearizable {:model m}) — no call site in the reference; it would be added to
 * the compose maps at set_full.clj:156-158 / tests/ledger.clj:363-367.
 * shards[n_shards] is caller-allocated.  Returns 0 on success (verdict in out), <0 on error
 * (jtb_last_error; glue throws so that jepsen's check-safe yields {:valid? :unknown}).          */
int jtb_check_linearizable(jtb_ctx* ctx, const jtb_history* h, const jtb_model* m,
                           jtb_lin_shard* shards, jtb_lin_result* out);

/* ---- knossos analysis :configs ------------------------------------------------ *
 * The configurations alive where an INVALID shard got stuck: every visited configuration whose first
 * un-linearized :ok return is the witness (`:op`).  knossos reports them as {:model :pending ...} maps and
 * jepsen.checker/linearizable keeps the first 10.  Call directly after a jtb_check_linearizable(ctx, h, m, ..)
 * that reported JTB_INVALID for `shard`, with the SAME h and m (the visited table of that search is read;
 * any other call on ctx invalidates it).  Writes min(cap, *n_total) configurations in a canonical order
 * (ascending, lexicographic over the struct's int32 fields in declaration order; unused entries are 0).
 * Returns 0, <0 on error. */
typedef struct jtb_final_config {
    int32_t state;                          /* register / cas-register value (JTB_NIL = nil); 0 for bank, set */
    int32_t balances[JTB_MAX_ACCOUNTS];     /* bank: balance per account slot                                  */
    int32_t n_pending;                      /* completed ops open at the witness' return, NOT linearized
                                               (the witness itself is one of them)                             */
    int32_t n_linearized_open;              /* completed ops open at the witness' return, already linearized   */
    int32_t n_crashed_linearized;           /* crashed (:info) ops linearized in this configuration            */
    int32_t pending_index[64];              /* :index of their invocations, ascending                          */
    int32_t linearized_open_index[64];
} jtb_final_config;
int jtb_final_configs(jtb_ctx* ctx, const jtb_history* h, const jtb_model* m, int32_t shard,
                      jtb_final_config* out, int32_t cap, int64_t* n_total);

/* ---- hot path A4: (checker/set-full {:linearizable? L}) at set_full.clj:157 ------------------ */
int jtb_check_set_full(jtb_ctx* ctx, const jtb_history* h, int linearizable, jtb_setfull_out* out);

/* ---- hot path A8: bank SI checker, tests/ledger.clj:154-192 (after ledger->bank) -------------- */
int jtb_check_bank_totals(jtb_ctx* ctx, const jtb_history* h, const jtb_model* accounts,
                          int64_t total_amount, jtb_bank_result* out);

/* ---- multi-GPU fan-out inside the library (SURVEY §8(b) `n_gpus`, §8(e)) ----------------------------------- *
 * What `independent/checker` (set_full.clj:155) does over JVM threads, done over the GPUs of one box for a host
 * that is a single process (a JVM through JNI): the shards (independent keys) of the history are partitioned over
 * `n_gpus` devices (longest-processing-time-first on events^2), every device checks its share through its own
 * context on its own host thread, and the per-shard (verdict, witness, previous-ok) vectors are merged with ONE
 * ncclAllReduce(ncclMax) over int32[3 * n_shards] (ncclCommInitAll inside the library; NCCL is dlopen'ed at
 * jtb_multi_create, libjtb_check.so itself has no link-time dependency on it).  No configuration ever crosses GPUs:
 * linearizability is local (Herlihy-Wing), so the verdict lattice 0 < 1 < 2 under MAX is all that has to travel.
 * A history with ONE shard runs on the first device (it does not shard: SURVEY §8(e) "replicas only").
 * opts->device is ignored (devices 0 .. n_gpus-1); n_gpus <= 0 means every visible device.                     */
typedef struct jtb_multi jtb_multi;
jtb_multi*  jtb_multi_create(const jtb_opts* opts, int n_gpus);   /* NULL on failure (see jtb_multi_create_error) */
const char* jtb_multi_create_error(void);                          /* why the last jtb_multi_create returned NULL   */
void        jtb_multi_destroy(jtb_multi* mg);
int         jtb_multi_n_gpus(const jtb_multi* mg);
const char* jtb_multi_last_error(const jtb_multi* mg);
/* same contract as jtb_check_linearizable; out->seconds_kernel = max over devices; device_of_shard (may be NULL)
 * receives the device each shard was checked on */
int jtb_multi_check_linearizable(jtb_multi* mg, const jtb_history* h, const jtb_model* m,
                                 jtb_lin_shard* shards, jtb_lin_result* out, int32_t* device_of_shard);
/* same contract as jtb_check_set_full; only the per-shard structs are merged (the optional per-element and
 * suspect-read detail of `out` must be unset: elem_capacity == suspect_capacity == missing_capacity == 0) */
int jtb_multi_check_set_full(jtb_multi* mg, const jtb_history* h, int linearizable, jtb_setfull_out* out,
                             int32_t* device_of_shard);

/* ---- diagnostics: counters of the last jtb_check_linearizable call ------------------------------ *
 * out[0..11] = configs, probes, expansions, ring tail, ring head, idle polls, max probe length,
 * table slots, grid CTAs, ring entries, search launches (pause/resume growth + 1), kernel microseconds,
 * out[12..14] = host->device bytes, device->host bytes, CUDA kernels launched,
 * out[15..18] = scout steps, scout configs, shards decided by a scout, scouts launched */
int jtb_get_stats(jtb_ctx* ctx, unsigned long long* out, int n);

/* SURVEY 8(f) N2 — the step before the checkers, on the device.
 * jtb_partition_by_key replaces jepsen.independent/subhistory (set_full.clj:155: independent/checker re-filters the
 * whole history once per key): ONE stable partition of the events by key.  event_key[i] = the key of event i (any
 * int64; nemesis / un-keyed events can carry a key of their own).  Out: order[n_events] = original position of the
 * i-th event of the partitioned history (events of one key keep their history order), key_ids[n_keys] ascending,
 * shard_off[n_keys + 1] = the CSR offsets of jtb_history; key_cap = capacity of key_ids (shard_off: key_cap + 1).
 * jtb_ledger_balances is ledger->bank's arithmetic (tests/ledger.clj:100-105): balance = credits-posted - debits-posted. */
int jtb_partition_by_key(jtb_ctx* ctx, int64_t n_events, const int64_t* event_key, int32_t* order, int64_t* shard_off,
                         int64_t* key_ids, int32_t key_cap, int32_t* n_keys);
int jtb_ledger_balances(jtb_ctx* ctx, int64_t n, const int64_t* credits_posted, const int64_t* debits_posted,
                        int32_t* balance);

/* Page-locked host memory for the flattened arrays (what a JNI shim wraps in a direct ByteBuffer, what the Python
 * mirror backs its numpy arrays with).  Not required: any host pointer works; page-locked ones are copied by DMA at the
 * PCIe rate instead of being staged through the driver (the id lists of 100k-op set-full histories are ~600 MB).
 * NULL when the allocation fails. */
void* jtb_host_alloc(size_t bytes);
void  jtb_host_free(void* p);

/* ---- diagnostics: host preparation only (pairing, slots, tables) — no device work; returns seconds
 * or a negative value on malformed input.  Lets callers see the host share of time-to-verdict. */
double jtb_prepare_seconds(const jtb_history* h, const jtb_model* m);
/* same, also reporting the layout chosen: info[0..3] = key bytes, slot lanes (32|64), max crashed-op classes per
 * shard, total completed ops (ranks) */
double jtb_prepare_info(const jtb_history* h, const jtb_model* m, long long info[4]);

/* ---- K2 in isolation: visited-table probe/insert microbenchmark (roofline evidence) ----------- *
 * Inserts n_keys pseudo-random 128-bit keys then probes them `rounds` times; returns device
 * seconds for insert and probe phases.  variant selects the probe path.          */
int jtb_table_bench(jtb_ctx* ctx, uint64_t n_keys, int variant, int rounds,
                    double* insert_seconds, double* probe_seconds, uint64_t* found);

/* ---- the memory system under a hash probe: random 16 B gathers as a function of the footprint ------------------ *
 * Every thread issues `iters` rounds of `in_flight` (1, 2, 4, 8, 16) independent random 16 B loads (the search
 * kernel's ld.global.cg.v2.u64 probe; wide = 2: both halves of the 32 B sector) over a table of table_bytes
 * (rounded down to a power of two), ctas_per_sm x 256 threads per SM.  Returns the best of `rounds` timings and the
 * number of 16 B-slot probes issued.  profiles/ holds the sweep 64 MiB .. 16 GiB.               */
int jtb_gather_bench(jtb_ctx* ctx, uint64_t table_bytes, int in_flight, int wide, uint32_t iters, int ctas_per_sm,
                     int rounds, double* seconds, uint64_t* n_probes);

#ifdef __cplusplus
}
#endif
#endif /* JTB_CHECK_H */
