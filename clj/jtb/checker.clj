(ns jtb.checker
  "Clojure glue for the B200 history checker (libjtb_check.so through libjtb_jni.so).

  UNCOMPILED HERE: this image has no JVM, Clojure or jni.h.  The tested twin of this file is
  jepsen_tigerbeetle_b200/{history,checker}.py, which flattens the same op shapes and calls the same
  C ABI through ctypes.  Keep the two in lock-step.

  Drop-in use in the reference (nurturenature/jepsen-tigerbeetle):

    ;; src/tigerbeetle/workloads/set_full.clj:155-158
    :checker (jtb/independent-checker
              (checker/compose
               {:set-full              (jtb/set-full {:linearizable? true})    ; was checker/set-full
                :linear                (jtb/linearizable {:model :set})        ; new, optional
                :read-all-invoked-adds (read-all-invoked-adds)}))

    ;; src/tigerbeetle/tests/ledger.clj:363-367
    :checker (checker/compose
              {:SI     (jtb/bank-checker checker-opts)                         ; was (checker checker-opts)
               :linear (jtb/linearizable {:model :bank})                       ; new
               ...})"
  (:require [jepsen.checker :as checker]
            [jepsen.independent :as independent]
            [knossos.op :as op])
  (:import (jtb Native)))

;; ---- opcodes (include/jtb_check.h) -----------------------------------------------------------
(def type-code {:invoke 0 :ok 1 :fail 2 :info 3})
(def f-code    {:read 0 :write 1 :cas 2 :add 3 :transfer 4})
(def NIL Integer/MIN_VALUE)
(def verdict   {0 true 1 :unknown 2 false})
(def cause     {0 nil 1 :table-full 2 :budget 3 :too-wide})

(defn- ledger->bank-op
  "tests/ledger.clj:89-114 for one op; returns nil for :l-t ops."
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
      :l-t nil)))

(defn flatten-history
  "history (vector of op maps) -> map of primitive arrays laid out as `struct jtb_history`.
  Client ops only ((int? process)); independent tuples become CSR shards sorted by key."
  [model history]
  (let [ops   (->> history
                   (filter (comp int? :process))
                   (keep (fn [op] (if (= :txn (:f op)) (ledger->bank-op op) op)))
                   (map (fn [op] (if (independent/tuple? (:value op))
                                   (assoc op ::key (key (:value op)) :value (val (:value op)))
                                   (assoc op ::key nil)))))
        keys  (->> ops (map ::key) distinct (sort-by #(or % Long/MIN_VALUE)) vec)
        by-k  (group-by ::key ops)
        ops   (vec (mapcat by-k keys))
        n     (count ops)
        type  (byte-array n) f (byte-array n) flags (byte-array n)
        proc  (int-array n) index (int-array n) time (long-array n)
        a     (int-array n) b (int-array n) c (int-array n)
        poff  (long-array n) plen (int-array n)
        payload (java.util.ArrayList.)]
    (dotimes [i n]
      (let [{:keys [type f value process index time final?] :as o} (nth ops i)
            put-payload! (fn [xs] (aset poff i (long (.size payload)))
                           (if (nil? xs)
                             (aset plen i (int -1))
                             (do (aset plen i (int (count xs)))
                                 (doseq [x xs] (.add payload (int x))))))]
        (aset ^bytes type i (byte (type-code (:type o))))
        (aset ^bytes f i (byte (f-code (:f o))))
        (aset ^bytes flags i (byte (if final? 1 0)))
        (aset proc i (int process)) (aset index i (int index)) (aset ^longs time i (long (:time o)))
        (put-payload! nil)
        (case [model (:f o)]
          ([:register :read] [:cas-register :read])   (aset a i (int (if (nil? value) NIL value)))
          ([:register :write] [:cas-register :write]) (aset a i (int value))
          [:cas-register :cas] (do (aset a i (int (first value))) (aset b i (int (second value))))
          [:set :add]   (aset a i (int value))
          [:set :read]  (put-payload! (when value (sort value)))
          [:bank :read] (put-payload! (when (and value (= :ok (:type o)))
                                        (mapcat (fn [[id bal]] [id (if (nil? bal) NIL bal)]) value)))
          [:bank :transfer] (do (aset a i (int (:amount value)))
                                (aset b i (int (or (:debit-acct value) (:from value))))
                                (aset c i (int (or (:credit-acct value) (:to value))))))))
    {:n n :type type :f f :flags flags :process proc :index index :time time :a a :b b :c c
     :payload-off poff :payload-len plen :payload (int-array payload)
     :shard-off (long-array (reductions + 0 (map (comp count by-k) keys)))
     :keys keys :key-ids (long-array (map #(or % -1) keys))}))

(defonce ^:private ctx (delay (Native/create 0)))   ; one context per JVM; calls on it are serialised

(defn- decode-configs
  "int[] from Native/finalConfigs -> [{:model state-or-balances :pending [{:index i}..] ...}]"
  [^ints xs]
  (for [off (range 1 (alength xs) 140)
        :let [rec  (vec (java.util.Arrays/copyOfRange xs (int off) (int (+ off 140))))
              np   (rec 9) nl (rec 10)]]
    {:model             {:register (rec 0) :balances (subvec rec 1 9)}
     :pending           (mapv #(hash-map :index %) (subvec rec 12 (+ 12 np)))
     :linearized-open   (mapv #(hash-map :index %) (subvec rec 76 (+ 76 nl)))
     :crashed-linearized (rec 11)}))

(defn linearizable
  "Replacement for (checker/linearizable {:model m}); m in #{:register :cas-register :set :bank}."
  [{:keys [model]}]
  (assert model "The linearizable checker requires a model")
  (reify checker/Checker
    (check [_ test history _opts]
      (let [h   (flatten-history model history)
            res (Native/checkLinearizable @ctx h (name model)
                                          (int-array (:accounts test (range 1 9)))
                                          (boolean (:negative-balances? test true)))]
        ;; res: long[] {valid witness previous-ok cause configs probes} per shard, shard-major
        (let [[valid witness prev cause-code configs] (take 5 res)]
          (cond-> {:valid? (verdict valid) :analyzer :wgl-gpu :configs-explored configs}
            (= 2 valid) (assoc :op {:index witness} :previous-ok (when (<= 0 prev) {:index prev})
                               ;; knossos' :configs, first 10 like jepsen.checker/linearizable keeps them:
                               ;; int[] {total, then 140 ints per config in jtb_final_config field order}
                               :configs (decode-configs (Native/finalConfigs @ctx h (name model) 0 10)))
            (= 1 valid) (assoc :cause (cause cause-code))))))))

(defn set-full
  "Replacement for (checker/set-full {:linearizable? true}) — workloads/set_full.clj:157."
  [{:keys [linearizable?]}]
  (reify checker/Checker
    (check [_ _test history _opts]
      (Native/checkSetFull @ctx (flatten-history :set history) (boolean linearizable?)))))

(defn bank-checker
  "Replacement for the ledger :SI checker — tests/ledger.clj:154-192."
  [{:keys [negative-balances?]}]
  (reify checker/Checker
    (check [_ test history _opts]
      (Native/checkBankTotals @ctx (flatten-history :bank history)
                              (int-array (:accounts test)) (long (:total-amount test))
                              (boolean negative-balances?)))))

(defn independent-checker
  "Like independent/checker, but hands ALL keys to the native side in one call when the inner checker
  is one of ours (the GPU owns the fan-out); otherwise defers to jepsen.independent."
  [inner]
  (independent/checker inner))
