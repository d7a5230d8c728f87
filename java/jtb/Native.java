package jtb;

/**
 * JNI surface of the B200 history checker: one static native method per C-ABI entry point of libjtb_check.so
 * (include/jtb_check.h), implemented by jni/jtb_jni.c (libjtb_jni.so).  Used by clj/jtb/checker.clj.
 *
 * <p>A flattened history travels as ONE {@code Object[14]} of primitive arrays laid out like {@code struct
 * jtb_history}: {@code [byte[] type, byte[] f, byte[] flags, int[] process, int[] index, long[] timeNs, int[] a,
 * int[] b, int[] c, long[] payloadOff, int[] payloadLen, int[] payload, long[] shardOff, long[] keyIds]}.
 * Results come back as flat {@code long[]} / {@code int[]} records (layouts below) that the Clojure side turns into
 * the checker result maps.  A native error is a {@link RuntimeException}, so jepsen's {@code check-safe} reports
 * {@code {:valid? :unknown :error ...}} exactly as it does for a throwing Clojure checker.
 */
public final class Native {
    static {
        System.loadLibrary("jtb_jni");
    }

    private Native() {}

    /** {@code jtb_opts.flags} bits (include/jtb_check.h). */
    public static final int OPT_NO_EAGER_READS = 1, OPT_NO_SCOUTS = 2, OPT_ENGINE_LEVEL = 4, OPT_ENGINE_WORKLIST = 8,
            OPT_NO_BEAM = 16;

    /** Number of CUDA devices ({@code jtb_device_count}). */
    public static native int deviceCount();

    /** {@code jtb_create}: one context = one device + stream + cached buffers; calls on it are serialised. */
    public static native long create(int device, int flags, long tableBytes, long maxConfigs, int timeBudgetMs);

    public static native void destroy(long ctx);

    /** {@code jtb_multi_create}: in-library fan-out over nGpus devices (0 = all) with one NCCL all-reduce(MAX). */
    public static native long multiCreate(int nGpus, int flags, long tableBytes, long maxConfigs, int timeBudgetMs);

    public static native void multiDestroy(long multi);

    /**
     * {@code jtb_check_linearizable} (multi == false, handle from {@link #create}) or
     * {@code jtb_multi_check_linearizable} (multi == true, handle from {@link #multiCreate}).
     *
     * @return {@code [valid, nFailures, configs, probes, kernelNs, totalNs, keyBytes, nShards]} followed by 7 longs
     *     per shard: {@code valid, witnessIndex, previousOkIndex, cause, configs, probes, device}
     */
    public static native long[] checkLinearizable(long handle, boolean multi, Object[] history, int modelKind,
                                                  int initValue, int[] accounts, int[] initBalances, boolean negativeOk);

    /**
     * {@code jtb_final_configs}: knossos' {@code :configs} of an INVALID shard; call directly after
     * {@link #checkLinearizable} (multi == false) on the same history.
     *
     * @return {@code [total]} followed by min(cap, total) records of 140 ints in {@code jtb_final_config} field order
     */
    public static native int[] finalConfigs(long ctx, Object[] history, int modelKind, int initValue, int[] accounts,
                                            int[] initBalances, boolean negativeOk, int shard, int cap);

    /**
     * {@code jtb_check_set_full} / {@code jtb_multi_check_set_full} (per-shard structs only when multi).
     *
     * @return {@code [valid, nFailures, raiaValid, nSuspect, kernelNs, totalNs, nShards, nElems]}, then 10 longs per
     *     shard ({@code valid, attempt, stable, lost, neverRead, stale, duplicated, suspectFinalReads,
     *     stableLatencyMaxMs, lostLatencyMaxMs}), then {@code elemOff[nShards + 1]}, then 4 longs per element
     *     ({@code id, outcome, latencyMs, dupCount}), then per suspect final read {@code shard, index, nMissing,
     *     missing ids...}
     */
    public static native long[] checkSetFull(long handle, boolean multi, Object[] history, boolean linearizable);

    /**
     * {@code jtb_check_bank_totals}.
     *
     * @return {@code [valid, referenceThrows, readCount, errorCount, firstErrorIndex, firstErrorType, count[5],
     *     firstIndex[5], lastIndex[5], worstIndex[5], lowestTotal, highestTotal, lowestIndex, highestIndex, kernelNs,
     *     totalNs]} (34 longs)
     */
    public static native long[] checkBankTotals(long ctx, Object[] history, int[] accounts, long totalAmount,
                                                boolean negativeOk);
}
