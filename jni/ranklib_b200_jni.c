/*
 * ranklib_b200_jni.c — the thin JNI shim between RankLib's Java façades and libranklib_b200.so.
 *
 * There is no JDK in this image (no jni.h, no javac — SURVEY.md F1), so no JVM has loaded this file yet.  It is kept
 * mechanical — every native method is one call into the C ABI of include/ranklib_b200.h plus array pinning — and it IS
 * compiled and executed by the test suite: tests/test_zz_jni_shim.py builds it against a mock <jni.h> (tests/jni_mock/)
 * and runs the call sequence of jni/java/.../B200LambdaMART.java through a mock JNIEnv.  Build, where a JDK exists:
 *
 *   gcc -shared -fPIC -DRLB_HAVE_JNI -I$JAVA_HOME/include -I$JAVA_HOME/include/linux -I../include \
 *       ranklib_b200_jni.c -L../ranklib_b200/csrc -lranklib_b200 -o libranklib_b200_jni.so
 *
 * Java side (jni/java/): class ciir.umass.edu.learning.tree.NativeBridge declares the
 * `native` methods below; B200LambdaMART overrides LambdaMART.init()/learn() (R/learning/tree/LambdaMART.java:68-272)
 * and delegates.  A non-zero status becomes RankLibError.create(msg)
 * (R/utilities/RankLibError.java:25-42), the error convention of the reference.
 */
#ifdef RLB_HAVE_JNI
#include <jni.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "ranklib_b200.h"

#define BRIDGE(name) Java_ciir_umass_edu_learning_tree_NativeBridge_##name

/* RankLibError.create(String) (R/utilities/RankLibError.java:25-42) + Throw.  A failed lookup leaves the JVM's own
 * NoClassDefFoundError / NoSuchMethodError pending, which is the right thing to surface. */
static void throw_message(JNIEnv* env, const char* text) {
    jclass cls = (*env)->FindClass(env, "ciir/umass/edu/utilities/RankLibError");
    if (!cls) return;
    jmethodID create = (*env)->GetStaticMethodID(env, cls, "create", "(Ljava/lang/String;)Lciir/umass/edu/utilities/RankLibError;");
    if (!create) return;
    jstring msg = (*env)->NewStringUTF(env, text ? text : "ranklib_b200: unknown error");
    if (!msg) return; /* OutOfMemoryError pending */
    jobject err = (*env)->CallStaticObjectMethod(env, cls, create, msg);
    if (err) (*env)->Throw(env, (jthrowable)err);
}

static void throw_ranklib_error(JNIEnv* env, rlb_ctx* ctx) { throw_message(env, rlb_last_error(ctx)); }

#define CHECK(ctx, call)                          \
    do {                                          \
        if ((call) != RLB_OK) {                   \
            throw_ranklib_error(env, (ctx));      \
            return 0;                             \
        }                                         \
    } while (0)

/* long create(int device) */
JNIEXPORT jlong JNICALL BRIDGE(create)(JNIEnv* env, jclass c, jint device) {
    rlb_ctx* ctx = NULL;
    CHECK(NULL, rlb_create(device, &ctx));
    return (jlong)(intptr_t)ctx;
}

JNIEXPORT jint JNICALL BRIDGE(destroy)(JNIEnv* env, jclass c, jlong h) { return rlb_destroy((rlb_ctx*)(intptr_t)h); }

/* int loadDense(long h, float[] X (N x F row-major, column j = features[j], NaN = unknown), long N, int F,
 *               int[] features, float[] labels, int[] qoff)
 * X is filled on the Java side by the loop that LambdaMART.init already runs over martSamples
 * (LambdaMART.java:82-91) calling DataPoint.getFeatureValue (DenseDataPoint.java:21-32). */
JNIEXPORT jint JNICALL BRIDGE(loadDense)(JNIEnv* env, jclass c, jlong h, jfloatArray X, jlong N, jint F, jintArray features,
                                          jfloatArray labels, jintArray qoff) {
    rlb_ctx* ctx = (rlb_ctx*)(intptr_t)h;
    jint Q = (*env)->GetArrayLength(env, qoff) - 1;
    /* the native side reads N*F values, N labels, F ids and Q+1 offsets: the Java arrays must hold them */
    if (N < 0 || F <= 0 || Q < 1 || (int64_t)(*env)->GetArrayLength(env, X) < (int64_t)N * F ||
        (int64_t)(*env)->GetArrayLength(env, labels) < (int64_t)N || (*env)->GetArrayLength(env, features) < F) {
        throw_message(env, "NativeBridge.loadDense: array lengths do not match N, F and the number of lists");
        return 0;
    }
    float* x = (*env)->GetPrimitiveArrayCritical(env, X, NULL);
    int32_t* f = (*env)->GetPrimitiveArrayCritical(env, features, NULL);
    float* l = (*env)->GetPrimitiveArrayCritical(env, labels, NULL);
    int32_t* q = (*env)->GetPrimitiveArrayCritical(env, qoff, NULL);
    int rc = rlb_load_dense(ctx, x, (int64_t)N, F, f, l, q, Q);
    (*env)->ReleasePrimitiveArrayCritical(env, qoff, q, JNI_ABORT);
    (*env)->ReleasePrimitiveArrayCritical(env, labels, l, JNI_ABORT);
    (*env)->ReleasePrimitiveArrayCritical(env, features, f, JNI_ABORT);
    (*env)->ReleasePrimitiveArrayCritical(env, X, x, JNI_ABORT);
    CHECK(ctx, rc);
    return 0;
}

/* int loadLetorFile(long h, String path, boolean mustHaveRelDoc, int[] features (null = all), int[] dims (out: N, Q, maxFid))
 * FeatureManager.readInput (R/features/FeatureManager.java:187-245) + the flattening of LambdaMART.init in one native step:
 * the file is parsed by the library's multithreaded reader and goes to the device without becoming DataPoint objects. */
JNIEXPORT jint JNICALL BRIDGE(loadLetorFile)(JNIEnv* env, jclass c, jlong h, jstring path, jboolean mustHaveRelDoc, jintArray features,
                                              jintArray dims) {
    rlb_ctx* ctx = (rlb_ctx*)(intptr_t)h;
    rlb_letor* set = NULL;
    const char* p = (*env)->GetStringUTFChars(env, path, NULL);
    int rc;
    if (!p) return 0; /* OutOfMemoryError is already pending */
    rc = rlb_letor_read(p, mustHaveRelDoc ? 1 : 0, 0, &set);
    (*env)->ReleaseStringUTFChars(env, path, p);
    if (rc != RLB_OK) {
        throw_ranklib_error(env, NULL); /* reader errors are context-free */
        return 0;
    }
    {
        int64_t n = 0;
        int32_t q = 0, mf = 0;
        jint d[3];
        rlb_letor_dims(set, &n, &q, &mf, NULL);
        d[0] = (jint)n;
        d[1] = q;
        d[2] = mf;
        if (features) {
            jsize nf = (*env)->GetArrayLength(env, features);
            jint* f = (*env)->GetIntArrayElements(env, features, NULL);
            rc = rlb_load_letor(ctx, set, (const int32_t*)f, nf);
            (*env)->ReleaseIntArrayElements(env, features, f, JNI_ABORT);
        } else {
            rc = rlb_load_letor(ctx, set, NULL, 0);
        }
        rlb_letor_free(set);
        CHECK(ctx, rc);
        (*env)->SetIntArrayRegion(env, dims, 0, 3, d);
    }
    return 0;
}

/* int init(long h, int nLeaves, int minLeafSupport, float learningRate, int nThreshold, int kind, int metric, int k,
 *          float featureSamplingRate, long seed) — the static fields of LambdaMART.java:37-42 at init() time */
JNIEXPORT jint JNICALL BRIDGE(init)(JNIEnv* env, jclass c, jlong h, jint nLeaves, jint mls, jfloat lr, jint nThreshold, jint kind,
                                     jint metric, jint k, jfloat frate, jlong seed) {
    rlb_ctx* ctx = (rlb_ctx*)(intptr_t)h;
    rlb_params p;
    p.n_leaves = nLeaves;
    p.min_leaf_support = mls;
    p.learning_rate = lr;
    p.n_threshold = nThreshold;
    p.kind = kind;
    p.metric = metric;
    p.metric_k = k;
    p.feature_sampling_rate = frate;
    p.seed = seed;
    CHECK(ctx, rlb_lambdamart_init(ctx, &p));
    return 0;
}

/* float boostIter(long h, int[] nodeInts (cap x 7), float[] nodeFloats (cap x 2), double[] nodeDeviance (cap), int[] nNodes)
 * One pass of the loop body LambdaMART.java:180-251; returns NDCG@k-T.  The Java side rebuilds the
 * Split objects (R/learning/tree/Split.java) from the flat arrays and calls ensemble.add(rt, learningRate). */
JNIEXPORT jfloat JNICALL BRIDGE(boostIter)(JNIEnv* env, jclass c, jlong h, jintArray nodeInts, jfloatArray nodeFloats,
                                            jdoubleArray nodeDeviance, jintArray nNodes) {
    rlb_ctx* ctx = (rlb_ctx*)(intptr_t)h;
    jint cap = (*env)->GetArrayLength(env, nodeDeviance);
    rlb_node* nodes;
    int32_t n = 0;
    float metric = 0.f;
    /* the three arrays must agree: cap nodes of 7 ints, 2 floats, 1 double.  Checked BEFORE the library runs: a boosting
     * iteration that has run cannot be taken back, and the ensemble on the Java side would be out of step with the device. */
    if (cap < 1 || (*env)->GetArrayLength(env, nodeInts) < 7 * cap || (*env)->GetArrayLength(env, nodeFloats) < 2 * cap ||
        (*env)->GetArrayLength(env, nNodes) < 1) {
        throw_message(env, "NativeBridge.boostIter: node arrays of inconsistent lengths (7 ints, 2 floats, 1 double per node)");
        return 0.f;
    }
    nodes = (rlb_node*)malloc(sizeof(rlb_node) * (size_t)cap);
    if (!nodes) {
        throw_message(env, "NativeBridge.boostIter: out of memory");
        return 0.f;
    }
    if (rlb_boost_iter(ctx, nodes, cap, &n, &metric) != RLB_OK) {
        free(nodes);
        throw_ranklib_error(env, ctx);
        return 0.f;
    }
    jint* ni = (*env)->GetPrimitiveArrayCritical(env, nodeInts, NULL);
    jfloat* nf = (*env)->GetPrimitiveArrayCritical(env, nodeFloats, NULL);
    jdouble* nd = (*env)->GetPrimitiveArrayCritical(env, nodeDeviance, NULL);
    for (int i = 0; i < n; i++) {
        ni[7 * i + 0] = nodes[i].feature_id;
        ni[7 * i + 1] = nodes[i].feature_idx;
        ni[7 * i + 2] = nodes[i].threshold_idx;
        ni[7 * i + 3] = nodes[i].left;
        ni[7 * i + 4] = nodes[i].right;
        ni[7 * i + 5] = nodes[i].count;
        ni[7 * i + 6] = 0;
        nf[2 * i + 0] = nodes[i].threshold;
        nf[2 * i + 1] = nodes[i].output;
        nd[i] = nodes[i].deviance;
    }
    (*env)->ReleasePrimitiveArrayCritical(env, nodeDeviance, nd, 0);
    (*env)->ReleasePrimitiveArrayCritical(env, nodeFloats, nf, 0);
    (*env)->ReleasePrimitiveArrayCritical(env, nodeInts, ni, 0);
    free(nodes);
    (*env)->SetIntArrayRegion(env, nNodes, 0, 1, (jint*)&n);
    return metric;
}

/* int readScores(long h, double[] out) — modelScores, for scorer.score(rank(samples)) after learn() */
JNIEXPORT jint JNICALL BRIDGE(readScores)(JNIEnv* env, jclass c, jlong h, jdoubleArray out) {
    rlb_ctx* ctx = (rlb_ctx*)(intptr_t)h;
    jsize n = (*env)->GetArrayLength(env, out);
    jdouble* o = (*env)->GetPrimitiveArrayCritical(env, out, NULL);
    int rc = rlb_read(ctx, RLB_READ_SCORE, o, (int64_t)n * 8);
    (*env)->ReleasePrimitiveArrayCritical(env, out, o, 0);
    CHECK(ctx, rc);
    return 0;
}

/* rlb_node array from the flat Java arrays (feature id, children, threshold, output); NULL + pending exception on
 * inconsistent lengths or exhausted memory */
static rlb_node* unpack_nodes(JNIEnv* env, jintArray nodeInts, jfloatArray nodeFloats, jintArray treeOff, jfloatArray weights,
                              jsize* nTreesOut) {
    jsize nTrees = (*env)->GetArrayLength(env, weights);
    jsize nNodes = (*env)->GetArrayLength(env, nodeInts) / 7;
    if ((*env)->GetArrayLength(env, nodeFloats) < 2 * nNodes || (*env)->GetArrayLength(env, treeOff) != nTrees + 1) {
        throw_message(env, "NativeBridge: model arrays of inconsistent lengths (7 ints + 2 floats per node, nTrees + 1 offsets)");
        return NULL;
    }
    rlb_node* nodes = (rlb_node*)malloc(sizeof(rlb_node) * (size_t)(nNodes > 0 ? nNodes : 1));
    if (!nodes) {
        throw_message(env, "NativeBridge: out of memory");
        return NULL;
    }
    jint* ni = (*env)->GetIntArrayElements(env, nodeInts, NULL);
    jfloat* nf = (*env)->GetFloatArrayElements(env, nodeFloats, NULL);
    if (!ni || !nf) {
        if (ni) (*env)->ReleaseIntArrayElements(env, nodeInts, ni, JNI_ABORT);
        if (nf) (*env)->ReleaseFloatArrayElements(env, nodeFloats, nf, JNI_ABORT);
        free(nodes);
        return NULL; /* OutOfMemoryError pending */
    }
    for (jsize i = 0; i < nNodes; i++) {
        memset(&nodes[i], 0, sizeof(rlb_node));
        nodes[i].feature_id = ni[7 * i + 0];
        nodes[i].left = ni[7 * i + 3];
        nodes[i].right = ni[7 * i + 4];
        nodes[i].threshold = nf[2 * i + 0];
        nodes[i].output = nf[2 * i + 1];
    }
    (*env)->ReleaseIntArrayElements(env, nodeInts, ni, JNI_ABORT);
    (*env)->ReleaseFloatArrayElements(env, nodeFloats, nf, JNI_ABORT);
    *nTreesOut = nTrees;
    return nodes;
}

/* int ensembleEval(long h, int[] nodeInts, float[] nodeFloats, int[] treeOff, float[] weights, float[] X, long N, int nCols, float[] out)
 * Ensemble.eval for a batch (R/learning/tree/Ensemble.java:110-116): Ranker.rank / Evaluator.score call this once per
 * file instead of once per DataPoint.  The library validates the model itself (child indices, offsets). */
JNIEXPORT jint JNICALL BRIDGE(ensembleEval)(JNIEnv* env, jclass c, jlong h, jintArray nodeInts, jfloatArray nodeFloats,
                                             jintArray treeOff, jfloatArray weights, jfloatArray X, jlong N, jint nCols,
                                             jfloatArray out) {
    rlb_ctx* ctx = (rlb_ctx*)(intptr_t)h;
    jsize nTrees = 0;
    if (N < 0 || nCols <= 0 || (int64_t)(*env)->GetArrayLength(env, X) < (int64_t)N * nCols ||
        (int64_t)(*env)->GetArrayLength(env, out) < (int64_t)N) {
        throw_message(env, "NativeBridge.ensembleEval: X / out shorter than N x nCols / N");
        return 0;
    }
    rlb_node* nodes = unpack_nodes(env, nodeInts, nodeFloats, treeOff, weights, &nTrees);
    if (!nodes) return 0;
    jint* to = (*env)->GetPrimitiveArrayCritical(env, treeOff, NULL);
    jfloat* w = (*env)->GetPrimitiveArrayCritical(env, weights, NULL);
    jfloat* x = (*env)->GetPrimitiveArrayCritical(env, X, NULL);
    jfloat* o = (*env)->GetPrimitiveArrayCritical(env, out, NULL);
    int rc = rlb_ensemble_eval(ctx, nodes, to, nTrees, w, x, (int64_t)N, nCols, o);
    (*env)->ReleasePrimitiveArrayCritical(env, out, o, 0);
    (*env)->ReleasePrimitiveArrayCritical(env, X, x, JNI_ABORT);
    (*env)->ReleasePrimitiveArrayCritical(env, weights, w, JNI_ABORT);
    (*env)->ReleasePrimitiveArrayCritical(env, treeOff, to, JNI_ABORT);
    free(nodes);
    CHECK(ctx, rc);
    return 0;
}

/* int loadValidation(long h, float[] X (N x F, the training set's feature columns), long N, int F, float[] labels, int[] qoff)
 * Ranker.setValidationSet + modelScoresOnValidation of LambdaMART.init (LambdaMART.java:152-158): the validation lists stay
 * on the device; every boostIter then scores the new tree on them (LambdaMART.java:228-237) without host traffic. */
JNIEXPORT jint JNICALL BRIDGE(loadValidation)(JNIEnv* env, jclass c, jlong h, jfloatArray X, jlong N, jint F, jfloatArray labels,
                                               jintArray qoff) {
    rlb_ctx* ctx = (rlb_ctx*)(intptr_t)h;
    jint Q = (*env)->GetArrayLength(env, qoff) - 1;
    if (N < 0 || F <= 0 || Q < 1 || (int64_t)(*env)->GetArrayLength(env, X) < (int64_t)N * F ||
        (int64_t)(*env)->GetArrayLength(env, labels) < (int64_t)N) {
        throw_message(env, "NativeBridge.loadValidation: array lengths do not match N, F and the number of lists");
        return 0;
    }
    float* x = (*env)->GetPrimitiveArrayCritical(env, X, NULL);
    float* l = (*env)->GetPrimitiveArrayCritical(env, labels, NULL);
    int32_t* q = (*env)->GetPrimitiveArrayCritical(env, qoff, NULL);
    int rc = rlb_load_validation(ctx, x, (int64_t)N, F, l, q, Q);
    (*env)->ReleasePrimitiveArrayCritical(env, qoff, q, JNI_ABORT);
    (*env)->ReleasePrimitiveArrayCritical(env, labels, l, JNI_ABORT);
    (*env)->ReleasePrimitiveArrayCritical(env, X, x, JNI_ABORT);
    CHECK(ctx, rc);
    return 0;
}

/* float validMetric(long h) — computeModelScoreOnValidation() of the iteration boostIter just ran (LambdaMART.java:236) */
JNIEXPORT jfloat JNICALL BRIDGE(validMetric)(JNIEnv* env, jclass c, jlong h) {
    rlb_ctx* ctx = (rlb_ctx*)(intptr_t)h;
    float v = 0.f;
    CHECK(ctx, rlb_valid_metric(ctx, &v));
    return v;
}

/* double scoreResident(long h, int which, int[] nodeInts, float[] nodeFloats, int[] treeOff, float[] weights)
 * scorer.score(rank(samples)) (LambdaMART.java:259,263) on the training (0) / validation (1) set resident on the device */
JNIEXPORT jdouble JNICALL BRIDGE(scoreResident)(JNIEnv* env, jclass c, jlong h, jint which, jintArray nodeInts, jfloatArray nodeFloats,
                                                 jintArray treeOff, jfloatArray weights) {
    rlb_ctx* ctx = (rlb_ctx*)(intptr_t)h;
    jsize nTrees = 0;
    double metric = 0.0;
    rlb_node* nodes = unpack_nodes(env, nodeInts, nodeFloats, treeOff, weights, &nTrees);
    if (!nodes) return 0.0;
    jint* to = (*env)->GetPrimitiveArrayCritical(env, treeOff, NULL);
    jfloat* w = (*env)->GetPrimitiveArrayCritical(env, weights, NULL);
    int rc = rlb_score_resident(ctx, which, nodes, to, nTrees, w, NULL, &metric);
    (*env)->ReleasePrimitiveArrayCritical(env, weights, w, JNI_ABORT);
    (*env)->ReleasePrimitiveArrayCritical(env, treeOff, to, JNI_ABORT);
    free(nodes);
    CHECK(ctx, rc);
    return metric;
}

/* int loadBag(long h, long src, int[] picks) — Sampler.doSampling on the device (RFRanker.java:80, Sampler.java:21-38): this
 * context becomes the bag of src's lists picks[0], picks[1], ... */
JNIEXPORT jint JNICALL BRIDGE(loadBag)(JNIEnv* env, jclass c, jlong h, jlong src, jintArray picks) {
    rlb_ctx* ctx = (rlb_ctx*)(intptr_t)h;
    jsize n = (*env)->GetArrayLength(env, picks);
    jint* p = (*env)->GetPrimitiveArrayCritical(env, picks, NULL);
    int rc = rlb_load_bag(ctx, (const rlb_ctx*)(intptr_t)src, (const int32_t*)p, n);
    (*env)->ReleasePrimitiveArrayCritical(env, picks, p, JNI_ABORT);
    CHECK(ctx, rc);
    return 0;
}

/* int commUniqueId(byte[] id128) / int commInit(long h, int rank, int world, byte[] id128): query-sharded training on the
 * GPUs of one box, one process (JVM) per GPU (include/ranklib_b200.h rlb_comm_*); rank 0 creates the id and ships it. */
JNIEXPORT jint JNICALL BRIDGE(commUniqueId)(JNIEnv* env, jclass c, jbyteArray id) {
    uint8_t buf[128];
    if ((*env)->GetArrayLength(env, id) < 128) {
        throw_message(env, "NativeBridge.commUniqueId: the id array must hold 128 bytes");
        return 0;
    }
    CHECK(NULL, rlb_comm_unique_id(buf));
    (*env)->SetByteArrayRegion(env, id, 0, 128, (const jbyte*)buf);
    return 0;
}

JNIEXPORT jint JNICALL BRIDGE(commInit)(JNIEnv* env, jclass c, jlong h, jint rank, jint world, jbyteArray id) {
    rlb_ctx* ctx = (rlb_ctx*)(intptr_t)h;
    uint8_t buf[128];
    if ((*env)->GetArrayLength(env, id) < 128) {
        throw_message(env, "NativeBridge.commInit: the id array must hold 128 bytes");
        return 0;
    }
    (*env)->GetByteArrayRegion(env, id, 0, 128, (jbyte*)buf);
    CHECK(ctx, rlb_comm_init(ctx, rank, world, buf));
    return 0;
}
#endif /* RLB_HAVE_JNI */
