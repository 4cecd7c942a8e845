/*
 * NativeBridge — the one class that declares the native methods of jni/ranklib_b200_jni.c.
 *
 * NOT COMPILED IN THIS IMAGE (no JDK, SURVEY.md F1).  The C side of every method below is compiled and
 * executed in tests/test_zz_jni_shim.py against a mock JNIEnv (tests/jni_mock/), so the argument order,
 * array layouts and the error path written here are the ones the shim implements.
 *
 * Node arrays ("flat tree", one row per rlb_node of include/ranklib_b200.h; node 0 is the root):
 *   nodeInts[7*i + 0] = RankLib feature id (-1: leaf)      nodeFloats[2*i + 0] = threshold
 *   nodeInts[7*i + 1] = index into features[]              nodeFloats[2*i + 1] = leaf output
 *   nodeInts[7*i + 2] = threshold index                    nodeDeviance[i]     = Split.deviance
 *   nodeInts[7*i + 3] = left child (-1: leaf)
 *   nodeInts[7*i + 4] = right child
 *   nodeInts[7*i + 5] = training samples in the node
 */
package ciir.umass.edu.learning.tree;

final class NativeBridge {
    static {
        System.loadLibrary("ranklib_b200_jni");
    }

    /* rlb_params.kind */
    static final int KIND_LAMBDAMART = 0;
    static final int KIND_MART = 1;

    /* rlb_params.metric (RLB_METRIC_* of include/ranklib_b200.h) */
    static final int METRIC_NDCG = 0;
    static final int METRIC_DCG = 1;
    static final int METRIC_ERR = 2;
    static final int METRIC_MAP = 3;
    static final int METRIC_PRECISION = 4;
    static final int METRIC_RR = 5;
    static final int METRIC_BEST = 6;

    /* capacity of the node arrays for a tree of nLeaves leaves: 2*nLeaves-1 nodes, but the root is split before the
     * leaf budget is looked at (RegressionTree.java:64-67), so nLeaves = 1 still yields 3 nodes */
    static int nodeCapacity(final int nLeaves) {
        return 2 * nLeaves + 1;
    }

    private NativeBridge() {
    }

    /* Every method throws ciir.umass.edu.utilities.RankLibError (created by the shim through
     * RankLibError.create(String)) when the library reports a non-zero status. */
    static native long create(int device);

    static native int destroy(long handle);

    static native int loadDense(long handle, float[] X, long N, int F, int[] features, float[] labels, int[] qoff);

    /** Parses a LETOR text / .gz / binary-cache file natively and uploads it; dims receives {N, Q, maxFid}. */
    static native int loadLetorFile(long handle, String path, boolean mustHaveRelDoc, int[] features, int[] dims);

    static native int init(long handle, int nLeaves, int minLeafSupport, float learningRate, int nThreshold, int kind,
            int metric, int k, float featureSamplingRate, long seed);

    static native float boostIter(long handle, int[] nodeInts, float[] nodeFloats, double[] nodeDeviance, int[] nNodes);

    static native int readScores(long handle, double[] out);

    static native int ensembleEval(long handle, int[] nodeInts, float[] nodeFloats, int[] treeOff, float[] weights,
            float[] X, long N, int nCols, float[] out);

    /**
     * Ranker.setValidationSet + modelScoresOnValidation of LambdaMART.init (LambdaMART.java:152-158): the validation lists,
     * in the training set's feature columns, stay on the device; every boostIter then updates their cached scores and
     * evaluates the metric there (LambdaMART.java:228-237).
     */
    static native int loadValidation(long handle, float[] X, long N, int F, float[] labels, int[] qoff);

    /** computeModelScoreOnValidation() (LambdaMART.java:485-518) of the iteration boostIter just ran. */
    static native float validMetric(long handle);

    /**
     * scorer.score(rank(samples)) (LambdaMART.java:259,263) on the training (which = 0) / validation (1) set that is
     * resident on the device, for the model given as flat node arrays.
     */
    static native double scoreResident(long handle, int which, int[] nodeInts, float[] nodeFloats, int[] treeOff,
            float[] weights);

    /**
     * Sampler.doSampling on the device (RFRanker.java:80, Sampler.java:21-38): the context `handle` becomes the bag of the
     * lists picks[0], picks[1], ... of the context `source`, which holds the whole training set (loadDense).
     */
    static native int loadBag(long handle, long source, int[] picks);

    /**
     * Query-sharded training on the GPUs of one box, one JVM per GPU: rank 0 calls commUniqueId and ships the 128 bytes
     * to the other ranks; every rank calls commInit before loadDense of ITS shard of the lists.
     */
    static native int commUniqueId(byte[] id128);

    static native int commInit(long handle, int rank, int world, byte[] id128);
}
