(ns jtb.checker
  "Clojure glue for the B200 history checker: jepsen.checker/Checker implementations that flatten the history, call
  libjtb_check.so through jtb.Native (java/jtb/Native.java -> jni/jtb_jni.c) and build the SAME result maps the
  reference's checkers return.

  This image has no JVM/Clojure: the file cannot be loaded here.  What IS checked here (tests/test_jni_shim.py):
  every Native/<method> call below exists in jtb/Native.java with that arity, every native method has its
  Java_jtb_Native_<method> export in jni/jtb_jni.c, and the shim itself runs end to end against the library through
  a fake JNIEnv.  The tested twin of the map-building below is jepsen_tigerbeetle_b200/{history,checker}.py.

  Drop-in use in the reference (nurturenature/jepsen-tigerbeetle):

    ;; src/tigerbeetle/workloads/set_full.clj:155-158
    :checker (jtb/independent-checker                                           ; was independent/checker + compose
              {:set-full              [:set-full {:linearizable? true}]         ; was checker/set-full
               :linear                [:linearizable {:model :set}]             ; new, optional
               :read-all-invoked-adds [:read-all-invoked-adds]})                ; was (read-all-invoked-adds)
    ;; or, keeping jepsen's own fan-out (one native call per key):
    :checker (independent/checker
              (checker/compose {:set-full              (jtb/set-full {:linearizable? true})
                                :read-all-invoked-adds (jtb/read-all-invoked-adds)}))

    ;; src/tigerbeetle/tests/ledger.clj:363-367
    :checker (checker/compose
              {:SI     (jtb/bank-checker checker-opts)                          ; was (checker checker-opts)
               :linear (jtb/linearizable {:model :bank})                        ; new
               ...})"
  (:require [jepsen.checker :as checker]
            [jepsen.independent :as independent])
  (:import (jtb Native)))

;; ---- codes (include/jtb_check.h) ---------------------------------------------------------------------------
(def type-code  {:invoke 0 :ok 1 :fail 2 :info 3})
(def f-code     {:read 0 :write 1 :cas 2 :add 3 :transfer 4})
(def model-code {:register 0 :cas-register 1 :set 2 :bank 3})
(def NIL Integer/MIN_VALUE)
(def verdict    {0 true 1 :unknown 2 false})
(def cause      {0 nil 1 :table-full 2 :budget 3 :too-wide})
(def bank-error {1 :unexpected-key 2 :nil-balance 3 :wrong-total 4 :negative-value})

(defn- ledger->bank-op
  "tests/ledger.clj:89-114 for one client op; nil for :l-t ops (dropped there too)."
  [{:keys [type value] :as op}]
  (let [[f _ _] (first value)]
    (case f
      :r   (if (= :ok type)
             (assoc op :f :read
                    :value (reduce (fn [m [_ id {:keys [debits-posted credits-posted]}]]
                                     (assoc m id (- credits-posted debits-posted)))
                                   {} value))
             (assoc op :f :read :value nil))
      :t   (let [[_ _ v] (first value)] (assoc op :f :transfer :value v))
      :l-t nil
      nil)))

(defn- int-or-nil [x] (if (nil? x) NIL (int x)))

(defn flatten-history
  "history (seq of op maps) -> {:arrays Object[14] laid out as `struct jtb_history`, :keys [k ...], :by-index {idx op}}.
  Client ops only ((int? process)); independent tuples become CSR shards sorted by key; an op whose :value is a
  [nil nil] tuple (set_full.clj:112-116 with a rejected account creation) is skipped, like an element that was never
  tracked."
  [model history]
  (let [ops   (->> history
                   (filter (comp int? :process))
                   (keep (fn [op] (if (= :txn (:f op)) (ledger->bank-op op) op)))
                   (keep (fn [op]
                           (let [v (:value op)]
                             (if (independent/tuple? v)
                               (when-not (nil? (key v)) (assoc op ::key (key v) :value (val v)))
                               (assoc op ::key nil))))))
        ks    (->> ops (map ::key) distinct (sort-by #(or % Long/MIN_VALUE)) vec)
        by-k  (group-by ::key ops)
        ops   (vec (mapcat by-k ks))
        n     (count ops)
        type  (byte-array n) f (byte-array n) flags (byte-array n)
        proc  (int-array n) index (int-array n) time (long-array n)
        a     (int-array n) b (int-array n) c (int-array n)
        poff  (long-array n) plen (int-array n)
        payload (java.util.ArrayList.)]
    (dotimes [i n]
      (let [o     (nth ops i)
            value (:value o)
            put-payload! (fn [xs]
                           (aset poff i (long (.size payload)))
                           (if (nil? xs)
                             (aset plen i (int -1))
                             (do (aset plen i (int (count xs)))
                                 (doseq [x xs] (.add payload (int x))))))]
        (aset ^bytes type i (byte (type-code (:type o))))
        (aset ^bytes f i (byte (get f-code (:f o) 0)))
        (aset ^bytes flags i (byte (if (:final? o) 1 0)))
        (aset proc i (int (:process o)))
        (aset index i (int (:index o)))
        (aset ^longs time i (long (or (:time o) 0)))
        (put-payload! nil)
        (case [model (:f o)]
          ([:register :read] [:cas-register :read])   (aset a i (int-or-nil value))
          ([:register :write] [:cas-register :write]) (aset a i (int-or-nil value))
          [:cas-register :cas] (do (aset a i (int-or-nil (first value))) (aset b i (int-or-nil (second value))))
          [:set :add]   (aset a i (int-or-nil value))
          [:set :read]  (put-payload! (when (and value (= :ok (:type o))) (sort value)))
          [:bank :read] (put-payload! (when (and value (= :ok (:type o)))
                                        (mapcat (fn [[id bal]] [id (int-or-nil bal)]) value)))
          [:bank :transfer] (do (aset a i (int (:amount value)))
                                (aset b i (int (or (:debit-acct value) (:from value))))
                                (aset c i (int (or (:credit-acct value) (:to value)))))
          nil)))                                    ; any other :f: opcode 0 with no value — ignored by every checker
    {:arrays   (object-array [type f flags proc index time a b c poff plen (int-array payload)
                              (long-array (reductions + 0 (map (comp count by-k) ks)))
                              (long-array (map #(if (integer? %) (long %) -1) ks))])
     :keys     ks
     :by-index (into {} (map (juxt :index identity)) history)}))

;; ---- contexts -----------------------------------------------------------------------------------------------
;; checker/compose and independent/checker call `check` from several threads; a context serialises its calls, so
;; one context per device is enough (jtb_ctx holds a mutex).  `*n-gpus*` > 1 selects the in-library fan-out
;; (jtb_multi_*: shards partitioned over the GPUs, one NCCL all-reduce(MAX) of the verdict vector).
(def ^:dynamic *n-gpus* 1)
;; jtb_opts.flags (include/jtb_check.h, Native/OPT_*): 0 = defaults — eager reads, engine chosen from the history (level
;; engine for exhaustive sweeps, work list + beam + scouts for histories with crashed ops).  Rebind before first use, e.g.
;; (bit-or Native/OPT_NO_EAGER_READS Native/OPT_ENGINE_LEVEL) to sweep exactly the configurations Knossos would visit.
(def ^:dynamic *flags* 0)
(defonce ^:private ctx   (delay (Native/create 0 (int *flags*) 0 0 0)))
(defonce ^:private multi (delay (Native/multiCreate 0 (int *flags*) 0 0 0)))
(defn- handle [] (if (> *n-gpus* 1) [@multi true] [@ctx false]))

(defn- model-args [model test]
  (let [accounts (vec (:accounts test (range 1 9)))]
    [(int (model-code model))
     (int-or-nil (:init-value test))
     (int-array accounts)
     (int-array (map #(get (:initial-balances test) % 0) accounts))
     (boolean (:negative-balances? test true))]))

;; ---- linearizable ---------------------------------------------------------------------------------------------
(defn- decode-configs
  "int[] from Native/finalConfigs -> [{:model .. :pending [{:index i} ..] ..}] (jtb_final_config is 140 ints)"
  [model accounts ^ints xs]
  (vec (for [off (range 1 (alength xs) 140)
             :let [rec (vec (java.util.Arrays/copyOfRange xs (int off) (int (+ off 140))))
                   np  (rec 9) nl (rec 10)]]
         {:model              (case model
                                :bank (zipmap accounts (subvec rec 1 9))
                                :set  nil
                                (let [s (rec 0)] (when-not (= s NIL) s)))
          :pending            (mapv #(hash-map :index %) (subvec rec 12 (+ 12 np)))
          :linearized-open    (mapv #(hash-map :index %) (subvec rec 76 (+ 76 nl)))
          :crashed-linearized (rec 11)})))

(defn- lin-shard-map
  "7 longs of one shard (valid witness previous-ok cause configs probes device) -> knossos-style analysis map"
  [by-index [valid witness prev cause-code configs _probes device]]
  (cond-> {:valid? (verdict valid) :analyzer :wgl-gpu :configs-explored configs :device device}
    (= 2 valid) (assoc :op (by-index witness) :previous-ok (when (<= 0 prev) (by-index prev)))
    (= 1 valid) (assoc :cause (cause cause-code))))

(defn- check-linearizable* [model test history]
  (let [{:keys [arrays keys by-index]} (flatten-history model history)
        [h multi?] (handle)
        [kind init accounts balances neg-ok] (model-args model test)
        res    (Native/checkLinearizable h multi? arrays kind init accounts balances neg-ok)
        shards (mapv #(lin-shard-map by-index %) (partition 7 (drop 8 res)))
        ;; knossos' :configs (first 10, like jepsen.checker/linearizable keeps them) for INVALID shards; the visited
        ;; table of the search is read, so: single context only, straight after the search, before any other call
        shards (if multi?
                 shards
                 (vec (map-indexed
                       (fn [s m]
                         (if (false? (:valid? m))
                           (let [xs (Native/finalConfigs h arrays kind init accounts balances neg-ok (int s) (int 10))]
                             (assoc m :configs (decode-configs model (vec accounts) xs) :configs-total (aget xs 0)))
                           m))
                       shards)))]
    {:keys keys :shards shards}))

(defn linearizable
  "Replacement for (checker/linearizable {:model m}); m in #{:register :cas-register :set :bank}."
  [{:keys [model]}]
  (assert model "The linearizable checker requires a model")
  (reify checker/Checker
    (check [_ test history _opts]
      (first (:shards (check-linearizable* model test history))))))

;; ---- set-full + read-all-invoked-adds -----------------------------------------------------------------------
(defn- quantiles [xs]
  (let [xs (vec (sort xs)) n (count xs)]
    (when (pos? n)
      (into {} (for [p [0 0.5 0.95 0.99 1]] [p (nth xs (min (dec n) (long (Math/floor (* n p)))))])))))

(defn- set-full-maps
  "long[] of Native/checkSetFull -> {:per-shard [set-full result map ...] :suspects {shard [[index missing] ...]}}"
  [^longs res]
  (let [ns      (aget res 6)
        shard   (fn [s] (vec (java.util.Arrays/copyOfRange res (int (+ 8 (* 10 s))) (int (+ 18 (* 10 s))))))
        eoff0   (+ 8 (* 10 ns))
        eoff    (fn [s] (aget res (int (+ eoff0 s))))
        elems0  (+ eoff0 ns 1)
        elem    (fn [e] (let [o (int (+ elems0 (* 4 e)))] [(aget res o) (aget res (+ o 1)) (aget res (+ o 2)) (aget res (+ o 3))]))
        n-elems (aget res 7)
        per     (vec
                 (for [s (range ns)]
                   (let [[valid attempt stable lost never stale dup] (shard s)
                         es      (map elem (range (eoff s) (eoff (inc s))))
                         by-out  (group-by second es)           ; 0 never-read, 1 stable, 2 lost
                         stale-e (->> (by-out 1) (filter #(pos? (nth % 2))) (sort-by #(- (nth % 2))))]
                     {:valid?            (verdict valid)
                      :attempt-count     attempt
                      :stable-count      stable
                      :lost-count        lost
                      :lost              (into (sorted-set) (map first (by-out 2)))
                      :never-read-count  never
                      :never-read        (into (sorted-set) (map first (by-out 0)))
                      :stale-count       stale
                      :stale             (into (sorted-set) (map first stale-e))
                      :worst-stale       (mapv (fn [[id _ lat]] {:element id :stable-latency lat}) (take 8 stale-e))
                      :stable-latencies  (quantiles (map #(nth % 2) (by-out 1)))
                      :lost-latencies    (quantiles (map #(nth % 2) (by-out 2)))
                      :duplicated-count  dup
                      :duplicated        (into (sorted-map) (keep (fn [[id _ _ d]] (when (> d 1) [id d])) es))})))
        sus0    (+ elems0 (* 4 n-elems))
        suspects (loop [k (int sus0) acc {}]
                   (if (>= k (alength res))
                     acc
                     (let [s (aget res k) idx (aget res (+ k 1)) nm (aget res (+ k 2))
                           miss (into (sorted-set) (java.util.Arrays/copyOfRange res (int (+ k 3)) (int (+ k 3 nm))))]
                       (recur (int (+ k 3 nm)) (update acc s (fnil conj []) [idx miss])))))]
    {:per-shard per :suspects suspects}))

(defn- check-set-full* [linearizable? history]
  (let [{:keys [arrays keys]} (flatten-history :set history)
        [h multi?] (handle)]
    (assoc (set-full-maps (Native/checkSetFull h multi? arrays (boolean linearizable?))) :keys keys)))

(defn set-full
  "Replacement for (checker/set-full {:linearizable? true}) — workloads/set_full.clj:157."
  [{:keys [linearizable?]}]
  (reify checker/Checker
    (check [_ _test history _opts]
      (first (:per-shard (check-set-full* linearizable? history))))))

(defn- raia-map [suspects]
  (if (seq suspects) {:valid? false :suspect-final-reads suspects} {:valid? true}))

(defn read-all-invoked-adds
  "Replacement for (read-all-invoked-adds) — workloads/set_full.clj:51-75; evaluated on the device in the same
  pass as set-full."
  []
  (reify checker/Checker
    (check [_ _test history _opts]
      (raia-map (get (:suspects (check-set-full* true history)) 0)))))

;; ---- bank -----------------------------------------------------------------------------------------------------
(defn bank-checker
  "Replacement for the ledger :SI checker — tests/ledger.clj:154-192 (same result map)."
  [{:keys [negative-balances?]}]
  (reify checker/Checker
    (check [_ test history _opts]
      (let [{:keys [arrays by-index]} (flatten-history :bank history)
            accts (set (:accounts test))
            total (long (:total-amount test 0))
            res   (Native/checkBankTotals @ctx arrays (int-array (:accounts test)) total (boolean negative-balances?))
            err   (fn [t idx]                          ; the error map check-op builds (tests/ledger.clj:127-152)
                    (let [op (first (filter #(= :read (:f %)) (keep ledger->bank-op [(by-index idx)])))
                          v  (:value op)]
                      (case (bank-error t)
                        :unexpected-key {:type :unexpected-key :unexpected (remove accts (clojure.core/keys v)) :op op}
                        :nil-balance    {:type :nil-balance :nils (into {} (remove val v)) :op op}
                        :wrong-total    {:type :wrong-total :total (reduce + (vals v)) :op op}
                        :negative-value {:type :negative-value :negative (filter neg? (vals v)) :op op})))
            at    (fn [base t] (aget res (int (+ base t))))     ; count 6.., first 11.., last 16.., worst 21..
            errors (into {}
                         (for [t [1 2 3 4] :when (pos? (at 6 t))]
                           [(bank-error t)
                            (merge {:count (at 6 t)
                                    :first (err t (at 11 t))
                                    :worst (err t (at 21 t))
                                    :last  (err t (at 16 t))}
                                   (when (= t 3)
                                     {:lowest (err t (aget res 28)) :highest (err t (aget res 29))}))]))
            out   {:valid?      (verdict (aget res 0))
                   :read-count  (aget res 2)
                   :error-count (aget res 3)
                   :first-error (when (pos? (aget res 3)) (err (aget res 5) (aget res 4)))
                   :errors      errors}]
        (if (pos? (aget res 1))
          ;; tests/ledger.clj:122-123: with :total-amount 0 err-badness divides by zero as soon as util/max-by
          ;; compares two :wrong-total errors -> the reference checker throws -> check-safe reports :unknown
          (assoc out :valid? :unknown
                 :error "java.lang.ArithmeticException: Divide by zero (err-badness, tests/ledger.clj:122-123)")
          out)))))

;; ---- independent ----------------------------------------------------------------------------------------------
(defn independent-checker
  "Like (independent/checker (checker/compose checkers)) for a map {name checker-kind} built from THIS namespace's
  constructors, but ALL keys go to the native side in ONE call per checker kind — the GPU (or, with *n-gpus* > 1,
  the 8 GPUs of the box) owns the fan-out instead of a JVM thread pool.  Returns the same shape:
  {:valid? .. :results {k {name result .. :valid? ..}} :failures [k ..]}.
  `checkers` maps names to one of [:set-full opts] [:read-all-invoked-adds] [:linearizable opts]."
  [checkers]
  (reify checker/Checker
    (check [_ test history _opts]
      (let [sf   (delay (check-set-full* (boolean (some (fn [[_ [k o]]] (and (= k :set-full) (:linearizable? o))) checkers))
                                         history))
            cols (into {}
                       (for [[nm [kind o]] checkers]
                         [nm (case kind
                               :set-full              (zipmap (:keys @sf) (:per-shard @sf))
                               :read-all-invoked-adds (zipmap (:keys @sf)
                                                              (map-indexed (fn [s _] (raia-map (get (:suspects @sf) s)))
                                                                           (:keys @sf)))
                               :linearizable          (let [r (check-linearizable* (:model o) test history)]
                                                        (zipmap (:keys r) (:shards r))))]))
            ks      (sort (distinct (mapcat clojure.core/keys (vals cols))))
            results (into (sorted-map)
                          (for [k ks]
                            (let [m (into {} (for [[nm col] cols] [nm (get col k {:valid? :unknown})]))]
                              [k (assoc m :valid? (checker/merge-valid (map :valid? (vals m))))])))
            failures (vec (for [[k r] results :when (not (true? (:valid? r)))] k))]
        {:valid?   (checker/merge-valid (map :valid? (vals results)))
         :results  results
         :failures failures}))))
