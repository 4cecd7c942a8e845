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

static void throw_ranklib_error(JNIEnv* env, rlb_ctx* ctx) {
    jclass cls = (*env)->FindClass(env, "ciir/umass/edu/utilities/RankLibError");
    if (!cls) return;
    jmethodID create = (*env)->GetStaticMethodID(env, cls, "create", "(Ljava/lang/String;)Lciir/umass/edu/utilities/RankLibError;");
    jstring msg = (*env)->NewStringUTF(env, rlb_last_error(ctx));
    jobject err = (*env)->CallStaticObjectMethod(env, cls, create, msg);
    if (err) (*env)->Throw(env, (jthrowable)err);
}

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
    /* the three arrays must agree: cap nodes of 7 ints, 2 floats, 1 double */
    if (cap < 1 || (*env)->GetArrayLength(env, nodeInts) < 7 * cap || (*env)->GetArrayLength(env, nodeFloats) < 2 * cap ||
        (*env)->GetArrayLength(env, nNodes) < 1)
        cap = 0;
    nodes = (rlb_node*)malloc(sizeof(rlb_node) * (size_t)(cap > 0 ? cap : 1));
    if (!nodes) return 0.f;
    if (rlb_boost_iter(ctx, nodes, cap, &n, &metric) != RLB_OK) { /* cap == 0 ends here: "node buffer too small" */
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

/* int ensembleEval(long h, int[] nodeInts, float[] nodeFloats, int[] treeOff, float[] weights, float[] X, long N, int nCols, float[] out)
 * Ensemble.eval for a batch (R/learning/tree/Ensemble.java:110-116): Ranker.rank / Evaluator.score call this once per
 * file instead of once per DataPoint. */
JNIEXPORT jint JNICALL BRIDGE(ensembleEval)(JNIEnv* env, jclass c, jlong h, jintArray nodeInts, jfloatArray nodeFloats,
                                             jintArray treeOff, jfloatArray weights, jfloatArray X, jlong N, jint nCols,
                                             jfloatArray out) {
    rlb_ctx* ctx = (rlb_ctx*)(intptr_t)h;
    jsize nTrees = (*env)->GetArrayLength(env, weights);
    jsize nNodes = (*env)->GetArrayLength(env, nodeInts) / 7;
    jint* ni = (*env)->GetIntArrayElements(env, nodeInts, NULL);
    jfloat* nf = (*env)->GetFloatArrayElements(env, nodeFloats, NULL);
    rlb_node* nodes = (rlb_node*)malloc(sizeof(rlb_node) * (size_t)(nNodes > 0 ? nNodes : 1));
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
#endif /* RLB_HAVE_JNI */
